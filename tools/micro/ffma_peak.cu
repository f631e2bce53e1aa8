// FP32 SIMT peak on this GPU: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), register-resident, 16 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 bb = make_float2(b, b * 1.0001f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(a, bb.x, acc[i].x); acc[i].y = fmaf(a, bb.y, acc[i].y); }
      else acc[i] = __ffma2_rn(make_float2(a, a), bb, acc[i]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  for (int blocks_per_sm : {1, 2, 4, 8}) for (int mode = 0; mode < 2; ++mode) {
    const int iters = 20000, grid = 148 * blocks_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, 256>>>(d, iters, 1.0001f, 0.9999f); else k<1><<<grid, 256>>>(d, iters, 1.0001f, 0.9999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 32 * iters * (double)grid * 256;
    printf("%s  %d CTA/SM (%d warps/SM): %.2f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", blocks_per_sm, blocks_per_sm * 8, ms, flops / ms / 1e9);
  }
  return 0;
}
