// Does shared-memory fragment traffic at the GEMM ratio (LDS.128 : FFMA2 = 1 : 8, 8x8 thread tile) cap the FFMA2 rate?
// Variants: 0 = registers only, 1 = a/b fragments re-loaded from shared memory every step (as an SGEMM inner loop does),
//           2 = 8x16 thread tile (1 : 10.7), 3 = same loads but results unused by the FMAs (port pressure only)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, const float* in) {
  __shared__ __align__(16) float As[16][132];
  __shared__ __align__(16) float Bs[16][260];
  for (int i = threadIdx.x; i < 16 * 132; i += 256) (&As[0][0])[i] = in[i & 4095];
  for (int i = threadIdx.x; i < 16 * 260; i += 256) (&Bs[0][0])[i] = in[(i + 77) & 4095];
  __syncthreads();
  constexpr int TN2 = (MODE == 2) ? 8 : 4;               // packed column pairs per thread: 8 or 16 columns
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float a[8]; float2 b[TN2]; float2 acc[8][TN2];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + i];
#pragma unroll
  for (int j = 0; j < TN2; ++j) b[j] = make_float2(in[threadIdx.x + 64 + j], in[threadIdx.x + 96 + j]);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN2; ++j) acc[i][j] = make_float2(0.f, 0.f);
  float4 sink = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      if (MODE != 0) {
        const int kr = (kk + it) & 15;                   // (row depends on the outer iteration: the loads cannot be hoisted)
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kr][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kr][64 + ty * 4]);
        float4 bv[TN2 / 2];
#pragma unroll
        for (int g = 0; g < TN2 / 2; ++g) bv[g] = *reinterpret_cast<const float4*>(&Bs[kr][g * 64 + tx * 4]);
        if (MODE == 3) { sink.x += a0.x + a1.y; sink.y += bv[0].x + bv[1].y; }
        else {
          a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
          for (int g = 0; g < TN2 / 2; ++g) { b[2 * g] = make_float2(bv[g].x, bv[g].y); b[2 * g + 1] = make_float2(bv[g].z, bv[g].w); }
        }
      }
#pragma unroll
      for (int j = 0; j < TN2; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
    }
  }
  float s = sink.x + sink.y;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN2; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(float* d, float* in, const char* name, int bps) {
  const int iters = 2000, grid = 148 * bps;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fl = 2.0 * 16 * 8 * ((MODE == 2) ? 16 : 8) * iters * (double)grid * 256;
  printf("%-44s %d CTA/SM: %.2f ms  %.1f TFLOP/s\n", name, bps, ms, fl / ms / 1e9);
}
int main() {
  float *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 8192 * 4); cudaMemset(in, 0, 8192 * 4);
  for (int bps : {1, 2}) {
    run<0>(d, in, "registers only (8x8)", bps);
    run<1>(d, in, "fragments from shared, 8x8 tile (1:8)", bps);
    run<3>(d, in, "same LDS, FMAs independent of them", bps);
  }
  return 0;
}
