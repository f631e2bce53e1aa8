// FFMA2 rate with a GEMM-like register pattern: 8 scalar a x 4 packed b -> 32 packed accumulators (no memory).
#include <cstdio>
#include <cuda_runtime.h>
template <int ORDER> __global__ void __launch_bounds__(256) k(float* out, int iters, const float* in) {
  float a[8]; float2 b[4]; float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + i * 256];
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = make_float2(in[threadIdx.x + 2048 + j * 512], in[threadIdx.x + 2304 + j * 512]);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
    if (ORDER == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
    }
    // rotate operands so the loop is not trivially invariant (cheap: 2 moves per 32 FFMA2)
    const float t = a[0]; a[0] = a[7]; a[7] = t;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 8192 * 4); cudaMemset(in, 0, 8192 * 4);
  for (int bps : {1, 2}) for (int order = 0; order < 2; ++order) {
    const int iters = 20000, grid = 148 * bps;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (order == 0) k<0><<<grid, 256>>>(d, iters, in); else k<1><<<grid, 256>>>(d, iters, in);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("order %s, %d CTA/SM: %.2f ms  %.1f TFLOP/s\n", order ? "j-outer" : "i-outer", bps, ms, 2.0 * 64 * iters * (double)grid * 256 / ms / 1e9);
  }
  return 0;
}
