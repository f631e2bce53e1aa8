// Thin runtime layer: element-wise launch, memset, launch counting.
// GPU build: real kernels.  DX_EMU build (tests only): serial CPU loops.
#pragma once
#include "dx_common.h"
#include <atomic>
#include <stdlib.h>

namespace dx {

extern std::atomic<long long> g_launches;  // kernels launched by this library (bench.py reports it); entry points may run on several host threads

#ifndef DX_EMU
// ---- programmatic dependent launch --------------------------------------------------------
// A step of the path is several hundred short DEPENDENT launches on one stream (about 900 at batch 128), and between
// two of them the GPU idles for the launch latency of the second.  Every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and calls pdl_wait() (griddepcontrol.wait) before its first
// access to global memory: the CTAs of a launch are scheduled as the CTAs of the kernel in front EXIT (no kernel
// triggers earlier), do their set-up (barrier init and TMEM allocation in the GEMMs) and block until that kernel has
// completed and its writes are visible.  Every thread of every grid waits, so completion stays transitive along the
// stream.  Measured on B200: batch-128 train step 9.12 -> 7.26 ms, 65536-patch step 166.0 -> 165.0 ms.  (An explicit
// griddepcontrol.launch_dependents at kernel entry was measured too: 7.76 ms at batch 128 and 3 % SLOWER at 65536 —
// early-resident CTAs of the next kernel get in the way of the running one — so no kernel triggers.)
// DX_NO_PDL=1 launches without the attribute (the wait is then a no-op).
DX_D DX_INLINE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool pdl_enabled() { static const bool on = getenv("DX_NO_PDL") == nullptr; return on; }
template <class... KA, class... A>
inline cudaError_t launch_k(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (pdl_enabled()) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  if (cluster > 1) { at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = (unsigned)cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n; }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(args)...);
}

template <class F>
__global__ void __launch_bounds__(256) k_foreach(F f, int64_t n) {
  pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
// Grid sized in whole waves of the 148 SMs (8 resident 256-thread CTAs each).
template <class F>
inline void foreach (dx_stream_t s, int64_t n, F f) {
  if (n <= 0) return;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  launch_k(k_foreach<F>, dim3((unsigned)blocks), dim3(256), 0, s, 1, f, n);
  ++g_launches;
}
inline void zero_async(dx_stream_t s, void* p, size_t bytes) {
  if (bytes) cudaMemsetAsync(p, 0, bytes, s);
}
inline void copy_async(dx_stream_t s, void* dst, const void* src, size_t bytes) {
  if (bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
}
// zero `rows` rows of `width` bytes, `pitch` bytes apart (a column block of a wider matrix)
inline void zero2d_async(dx_stream_t s, void* p, size_t pitch, size_t width, size_t rows) {
  if (width && rows) cudaMemset2DAsync(p, pitch, 0, width, rows, s);
}
// Row compaction: rows[0..n) = ascending indices i < B with flag[i] != 0; *count = n.  One block scans the flags
// in 1024-wide chunks (ballot + popc inside warps, warp totals through shared memory).
static __global__ void __launch_bounds__(1024) k_compact_flags(int B, const uint8_t* __restrict__ flag, int* __restrict__ rows,
                                                               int* __restrict__ count) {
  pdl_wait();
  __shared__ int wsum[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < B; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const bool f = i < B && flag[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wsum[wid] = __popc(bal);
    __syncthreads();
    int off = base, tot = 0;
    for (int k = 0; k < 32; ++k) { if (k < wid) off += wsum[k]; tot += wsum[k]; }
    if (f) rows[off + __popc(bal & ((1u << lane) - 1))] = i;
    __syncthreads();
    if (threadIdx.x == 0) base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}
// Compacts and returns the count on the host (synchronises the stream: the caller sizes its next launches with it).
inline int compact_flags(dx_stream_t s, int B, const uint8_t* flag, int* rows, int* count_dev) {
  launch_k(k_compact_flags, dim3(1), dim3(1024), 0, s, 1, B, flag, rows, count_dev);
  ++g_launches;
  int n = 0;
  cudaMemcpyAsync(&n, count_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  return n;
}
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: CUDA error %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}
#else
template <class F>
inline void foreach (dx_stream_t, int64_t n, F f) {
  for (int64_t i = 0; i < n; ++i) f(i);
  ++g_launches;
}
inline void zero_async(dx_stream_t, void* p, size_t bytes) { if (bytes) memset(p, 0, bytes); }
inline void copy_async(dx_stream_t, void* dst, const void* src, size_t bytes) { if (bytes) memcpy(dst, src, bytes); }
inline void zero2d_async(dx_stream_t, void* p, size_t pitch, size_t width, size_t rows) {
  for (size_t r = 0; r < rows; ++r) memset((char*)p + r * pitch, 0, width);
}
inline int compact_flags(dx_stream_t, int B, const uint8_t* flag, int* rows, int* count_dev) {
  int n = 0;
  for (int i = 0; i < B; ++i) if (flag[i]) rows[n++] = i;
  *count_dev = n;
  ++g_launches;
  return n;
}
inline int check_launch(const char*) { return 0; }
#endif

// ---- math shared by host-emulation and device ------------------------------------
DX_HD DX_INLINE float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
DX_HD DX_INLINE float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }  // nn.Softplus(beta=1, threshold=20)

// ---- bump allocator over the caller's workspace -----------------------------------
struct Arena {
  char* base; size_t cap; size_t off; bool overflow;
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), overflow(false) {}
  template <class T> T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (off + bytes > cap) { overflow = true; off += bytes; return (T*)base; }
    T* r = (T*)(base + off); off += bytes; return r;
  }
};

}  // namespace dx
