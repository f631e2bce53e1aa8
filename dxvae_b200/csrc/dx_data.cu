// Batcher and data-format kernels.
//   batch_build_host / batch_schedule : flat CSR + topological level schedule with
//       feedback edges marked (replaces the per-graph DGL queries of model.py:164-177)
//   pack_graphs / unpack_graphs       : graph-major dxdata.py tensors <-> node-major
//   voices_to_graphs                  : dxdata.py:174-312 (_make_graph) on the device
//   pack_syx                          : dxdata.py:341-397 (graph_to_syx) voice packing
//   adamw_step, reparameterize        : model.py:375/386, :284
// All of it is integer/byte work bound by HBM traffic: one thread per output row,
// coalesced writes, no shared-memory staging needed at these sizes.
#include "dx_engine.h"
#include "dx_tables.h"

namespace dx {

// DX_ALGO (dxdata.py:140-171) as adjacency masks, bit (src*7+dst).
DX_HD DX_INLINE uint64_t algo_mask(int alg) {
  const uint64_t t[32] = {
      0x1808080208080ull, 0x808080218080ull, 0x1808010808080ull, 0x808410808080ull,
      0x1800880208080ull, 0x820880208080ull, 0x1804080208080ull, 0x804180208080ull,
      0x804080218080ull, 0x408011808080ull, 0x1408010808080ull, 0x204080218080ull,
      0x1204080208080ull, 0x1408080208080ull, 0x408080218080ull, 0x1801080408080ull,
      0x801080418080ull, 0x808021408080ull, 0x1c00810808080ull, 0x408011c04080ull,
      0xc00811c04080ull, 0x1e00810208080ull, 0x1c00810804080ull, 0x1e00810204080ull,
      0x1c00810204080ull, 0x1408010804080ull, 0x408011804080ull, 0x058080208080ull,
      0x1800880204080ull, 0x058080204080ull, 0x1800810204080ull, 0x1040810204080ull};
  return t[alg & 31];
}

// encode level of every node of one graph: 0 if no adjacent x>v, else 1+max level(x)
DX_HD DX_INLINE void graph_levels(uint64_t A, uint8_t* lv) {
  for (int v = NN - 1; v >= 0; --v) {
    int best = -1;
    for (int x = v + 1; x < NN; ++x)
      if (abit(A, x, v) | abit(A, v, x)) best = best > (int)lv[x] ? best : (int)lv[x];
    lv[v] = (uint8_t)(best + 1);
  }
}

// A node that receives an edge from a LOWER node (a feedback back-edge arrives at it): the only rows whose "out"
// projection half the encoder reads.  Inside a level those rows come first, so the half is a dense row prefix.
DX_HD DX_INLINE bool has_back_in(uint64_t A, int v) {
  for (int u = 0; u < v; ++u)
    if (abit(A, u, v)) return true;
  return false;
}

// level_ptr: 16 ints.  [0..7] level offsets (as before); [8 + L] = number of leading rows of level L with has_back_in.
static void schedule_host(int64_t B, const uint64_t* adj, uint8_t* level, int32_t* level_ptr, int32_t* level_rows,
                          int32_t* n_levels) {
  int maxl = 0;
  for (int64_t b = 0; b < B; ++b) {
    uint8_t lv[NN];
    graph_levels(adj[b], lv);
    for (int v = 0; v < NN; ++v) { level[b * NN + v] = lv[v]; if (v >= 1 && lv[v] > maxl) maxl = lv[v]; }
  }
  const int nl = B ? maxl + 1 : 0;
  int64_t pos = 0;
  for (int L = 0; L < 7; ++L) {
    level_ptr[L] = (int32_t)pos;
    level_ptr[8 + L] = 0;
    if (L < nl)
      for (int pass = 0; pass < 2; ++pass) {               // pass 0: rows a back-edge arrives at, pass 1: the rest
        for (int v = 1; v < NN; ++v)
          for (int64_t b = 0; b < B; ++b)
            if (level[b * NN + v] == L && has_back_in(adj[b], v) == (pass == 0)) level_rows[pos++] = (int32_t)(v * B + b);
        if (pass == 0) level_ptr[8 + L] = (int32_t)pos - level_ptr[L];
      }
  }
  level_ptr[7] = (int32_t)pos; level_ptr[15] = 0;
  *n_levels = nl;
}

int batch_build_host(int64_t B, const int32_t* edge_ptr, const int8_t* src, const int8_t* dst, uint64_t* adj,
                     int32_t* indptr, int32_t* indices, uint8_t* eflags, uint8_t* level, int32_t* level_ptr,
                     int32_t* level_rows, int32_t* n_levels) {
  DX_CHECK(B >= 0 && (int64_t)7 * B < (1ll << 31), "batch_build_host: batch too large");
  int64_t e = 0;
  indptr[0] = 0;
  for (int64_t b = 0; b < B; ++b) {
    uint64_t m = 0;
    for (int32_t k = edge_ptr[b]; k < edge_ptr[b + 1]; ++k) {
      const int s = src[k], d = dst[k];
      DX_CHECK(s >= 0 && s < NN && d >= 0 && d < NN, "batch_build_host: graph %lld has node id outside 0..6",
               (long long)b);
      m |= 1ull << (s * 7 + d);
    }
    adj[b] = m;
    // CSR by destination, sources ascending; duplicates collapse through the mask
    for (int d = 0; d < NN; ++d) {
      for (int s = 0; s < NN; ++s)
        if ((m >> (s * 7 + d)) & 1ull) {
          indices[e] = (int32_t)(b * NN + s);
          eflags[e] = (uint8_t)(s == d ? 2 : (s > d ? 0 : 1));
          ++e;
        }
      indptr[b * NN + d + 1] = (int32_t)e;
    }
  }
  schedule_host(B, adj, level, level_ptr, level_rows, n_levels);
  return 0;
}

#ifndef DX_EMU
namespace {
constexpr int SCH_T = 1024;  // graphs per block
constexpr int NBIN = 72;     // (level 0..5) x (back-edge target first, others second) x (operator 1..6), in that order

__global__ void __launch_bounds__(SCH_T) k_levels_count(int64_t B, const uint64_t* __restrict__ adj,
                                                        uint8_t* __restrict__ level, int32_t* __restrict__ counts,
                                                        int nblk) {
  pdl_wait();
  __shared__ int cnt[NBIN];
  if (threadIdx.x < NBIN) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * SCH_T + threadIdx.x;
  if (b < B) {
    uint8_t lv[NN];
    graph_levels(adj[b], lv);
    for (int v = 0; v < NN; ++v) level[b * NN + v] = lv[v];
    for (int v = 1; v < NN; ++v) atomicAdd(&cnt[lv[v] * 12 + (has_back_in(adj[b], v) ? 0 : 6) + (v - 1)], 1);
  }
  __syncthreads();
  if (threadIdx.x < NBIN) counts[threadIdx.x * nblk + blockIdx.x] = cnt[threadIdx.x];
}
// exclusive scan of counts in (bin, block) order -> offsets; level_ptr[L] = start of bin (L,1)
__global__ void k_scan_bins(int32_t* __restrict__ counts, int nblk, int32_t* __restrict__ level_ptr) {
  pdl_wait();
  __shared__ int tot[NBIN];
  const int bin = threadIdx.x;
  if (bin < NBIN) {
    int s = 0;
    for (int k = 0; k < nblk; ++k) { const int c = counts[bin * nblk + k]; counts[bin * nblk + k] = s; s += c; }
    tot[bin] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int q = 0; q < NBIN; ++q) {
      const int c = tot[q]; tot[q] = s;
      if (q % 12 == 0) level_ptr[q / 12] = s;
      if (q % 12 == 6) level_ptr[8 + q / 12] = s - level_ptr[q / 12];     // rows of this level a back-edge arrives at
      s += c;
    }
    level_ptr[6] = s; level_ptr[7] = s; level_ptr[14] = 0; level_ptr[15] = 0;
  }
  __syncthreads();
  if (bin < NBIN)
    for (int k = 0; k < nblk; ++k) counts[bin * nblk + k] += tot[bin];
}
__global__ void __launch_bounds__(SCH_T) k_scatter_rows(int64_t B, const uint8_t* __restrict__ level,
                                                        const uint64_t* __restrict__ adj,
                                                        const int32_t* __restrict__ offsets, int nblk,
                                                        int32_t* __restrict__ level_rows) {
  pdl_wait();
  __shared__ int wsum[32];
  const int64_t b = (int64_t)blockIdx.x * SCH_T + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint8_t lv[NN];
  for (int v = 0; v < NN; ++v) lv[v] = (b < B) ? level[b * NN + v] : 255;
  const uint64_t A = (b < B) ? adj[b] : 0ull;
  unsigned back = 0;
  for (int v = 1; v < NN; ++v) back |= (has_back_in(A, v) ? 1u : 0u) << v;
  for (int bin = 0; bin < NBIN; ++bin) {
    const int L = bin / 12, v = bin % 6 + 1;
    const bool f = lv[v] == L && (((back >> v) & 1u) != 0) == (bin % 12 < 6);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int inwarp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[wid] = __popc(bal);
    __syncthreads();
    int base = 0;
    for (int k = 0; k < wid; ++k) base += wsum[k];
    if (f) level_rows[offsets[bin * nblk + blockIdx.x] + base + inwarp] = (int32_t)(v * B + b);
    __syncthreads();
  }
}
}  // namespace

int batch_schedule(dx_stream_t st, int64_t B, const uint64_t* adj, uint8_t* level, int32_t* level_ptr,
                   int32_t* level_rows, int32_t* level_ptr_host, void* ws, size_t ws_bytes) {
  DX_CHECK(B > 0 && (int64_t)7 * B < (1ll << 31), "batch_schedule: bad batch size");
  const int nblk = (int)((B + SCH_T - 1) / SCH_T);
  Arena ar(ws, ws_bytes);
  int32_t* counts = ar.take<int32_t>((size_t)NBIN * nblk);
  DX_CHECK(!ar.overflow, "batch_schedule: workspace too small (%zu < %zu)", ws_bytes, ar.off);
  launch_k(k_levels_count, dim3(nblk), dim3(SCH_T), 0, st, 1, B, adj, level, counts, nblk);
  launch_k(k_scan_bins, dim3(1), dim3(96), 0, st, 1, counts, nblk, level_ptr);
  launch_k(k_scatter_rows, dim3(nblk), dim3(SCH_T), 0, st, 1, B, level, adj, counts, nblk, level_rows);
  g_launches += 3;
  cudaMemcpyAsync(level_ptr_host, level_ptr, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  return check_launch("batch_schedule");
}
#else
int batch_schedule(dx_stream_t, int64_t B, const uint64_t* adj, uint8_t* level, int32_t* level_ptr, int32_t* level_rows,
                   int32_t* level_ptr_host, void*, size_t) {
  int32_t nl = 0;
  schedule_host(B, adj, level, level_ptr, level_rows, &nl);
  memcpy(level_ptr_host, level_ptr, 16 * sizeof(int32_t));
  return 0;
}
#endif

// ---- decoder step schedule ---------------------------------------------------------------------
// Teacher forcing replays 21 (vi,vj) re-propagates per graph, but a re-propagate only changes the
// node state when the step adds an edge (vj->vi or vi->vj); otherwise it is the identity
// (model.py:353-358 with nothing added).  step t = vi*(vi-1)/2 + (vi-1-vj) (the order of
// model.py:311,347).  step_rows[step_ptr[t] .. step_ptr[t+1]) = ascending graph ids that are
// active at step t.
DX_HD DX_INLINE int step_vi(int t) { int vi = 1; while ((vi + 1) * vi / 2 <= t) ++vi; return vi; }
DX_HD DX_INLINE bool step_active(uint64_t A, int t) {
  if (t >= 27) {                                        // lists 27..32: node x = 0..5 has an edge to a HIGHER node
    const int x = t - 27;                                //   (a feedback back-edge leaves x: the decoder then needs the
    for (int v = x + 1; v < NN; ++v)                     //    "in" half of x's projections)
      if (abit(A, x, v)) return true;
    return false;
  }
  if (t >= 21) return abit(A, t - 20, t - 20) != 0;     // lists 21..26: self-loop on node 1..6
  const int vi = step_vi(t), vj = vi - 1 - (t - vi * (vi - 1) / 2);
  return (abit(A, vj, vi) | abit(A, vi, vj)) != 0;
}

static void steps_host(int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows) {
  int64_t pos = 0;
  for (int t = 0; t < NLIST; ++t) {
    step_ptr[t] = (int32_t)pos;
    for (int64_t b = 0; b < B; ++b)
      if (step_active(adj[b], t)) step_rows[pos++] = (int32_t)b;
  }
  step_ptr[NLIST] = (int32_t)pos;
}

#ifndef DX_EMU
namespace {
constexpr int NSTEPB = NLIST;
__global__ void __launch_bounds__(SCH_T) k_steps_count(int64_t B, const uint64_t* __restrict__ adj,
                                                       int32_t* __restrict__ counts, int nblk) {
  pdl_wait();
  __shared__ int cnt[NSTEPB];
  if (threadIdx.x < NSTEPB) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * SCH_T + threadIdx.x;
  if (b < B) {
    const uint64_t A = adj[b];
    for (int t = 0; t < NSTEPB; ++t) if (step_active(A, t)) atomicAdd(&cnt[t], 1);
  }
  __syncthreads();
  if (threadIdx.x < NSTEPB) counts[threadIdx.x * nblk + blockIdx.x] = cnt[threadIdx.x];
}
__global__ void k_steps_scan(int32_t* __restrict__ counts, int nblk, int32_t* __restrict__ step_ptr) {
  pdl_wait();
  __shared__ int tot[NSTEPB];
  const int bin = threadIdx.x;
  if (bin < NSTEPB) {
    int s = 0;
    for (int k = 0; k < nblk; ++k) { const int c = counts[bin * nblk + k]; counts[bin * nblk + k] = s; s += c; }
    tot[bin] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int q = 0; q < NSTEPB; ++q) { const int c = tot[q]; tot[q] = s; step_ptr[q] = s; s += c; }
    step_ptr[NSTEPB] = s;
  }
  __syncthreads();
  if (bin < NSTEPB)
    for (int k = 0; k < nblk; ++k) counts[bin * nblk + k] += tot[bin];
}
__global__ void __launch_bounds__(SCH_T) k_steps_scatter(int64_t B, const uint64_t* __restrict__ adj,
                                                         const int32_t* __restrict__ offsets, int nblk,
                                                         int32_t* __restrict__ step_rows) {
  pdl_wait();
  __shared__ int wsum[32];
  const int64_t b = (int64_t)blockIdx.x * SCH_T + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint64_t A = (b < B) ? adj[b] : 0ull;
  for (int t = 0; t < NSTEPB; ++t) {
    const bool f = (b < B) && step_active(A, t);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int inwarp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[wid] = __popc(bal);
    __syncthreads();
    int base = 0;
    for (int k = 0; k < wid; ++k) base += wsum[k];
    if (f) step_rows[offsets[t * nblk + blockIdx.x] + base + inwarp] = (int32_t)b;
    __syncthreads();
  }
}
}  // namespace

int batch_steps(dx_stream_t st, int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows,
                int32_t* step_ptr_host, void* ws, size_t ws_bytes) {
  DX_CHECK(B > 0 && (int64_t)NLIST * B < (1ll << 31), "batch_steps: bad batch size");
  const int nblk = (int)((B + SCH_T - 1) / SCH_T);
  Arena ar(ws, ws_bytes);
  int32_t* counts = ar.take<int32_t>((size_t)NSTEPB * nblk);
  DX_CHECK(!ar.overflow, "batch_steps: workspace too small (%zu < %zu)", ws_bytes, ar.off);
  launch_k(k_steps_count, dim3(nblk), dim3(SCH_T), 0, st, 1, B, adj, counts, nblk);
  launch_k(k_steps_scan, dim3(1), dim3(64), 0, st, 1, counts, nblk, step_ptr);
  launch_k(k_steps_scatter, dim3(nblk), dim3(SCH_T), 0, st, 1, B, adj, counts, nblk, step_rows);
  g_launches += 3;
  cudaMemcpyAsync(step_ptr_host, step_ptr, (NLIST + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  return check_launch("batch_steps");
}
#else
int batch_steps(dx_stream_t, int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows,
                int32_t* step_ptr_host, void*, size_t) {
  steps_host(B, adj, step_ptr, step_rows);
  memcpy(step_ptr_host, step_ptr, (NLIST + 1) * sizeof(int32_t));
  return 0;
}
#endif
int batch_steps_host(int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows) {
  steps_host(B, adj, step_ptr, step_rows);
  return 0;
}

// ---- layout conversion ---------------------------------------------------------------------
int pack_graphs(dx_stream_t st, int64_t B, const float* Xg, const float* Pg, float* Xn, int32_t* cls) {
  foreach (st, B * NN * XP, [=] DX_HD(int64_t i) {
    const int c = (int)(i % XP); const int64_t r = i / XP; const int64_t b = r % B; const int v = (int)(r / B);
    Xn[i] = c < SX ? Xg[(b * NN + v) * SX + c] : 0.f;
  });
  if (cls)
    foreach (st, B * 14, [=] DX_HD(int64_t i) {
      const int64_t b = i % B; const int k = (int)(i / B);
      float val;
      if (k == 0) val = Pg[(b * NN) * NP + 17];
      else if (k == 1) val = Pg[(b * NN) * NP + 18];
      else if (k < 8) val = Pg[(b * NN + (k - 1)) * NP + 19];
      else val = Pg[(b * NN + (k - 7)) * NP + 20];
      cls[i] = (int32_t)val;
    });
  return check_launch("pack_graphs");
}

// The same for the rows idx[b] of a larger graph-major store (a training batch drawn from a dataset).  Xg / Pg / adjg may
// be PINNED HOST memory (mapped into the device's address space): the gather then is the host-to-device transfer of the
// batch — 1.35 KB per graph read over the bus straight into the kernels' layout, no host-side gather, no staging copy.
int pack_graphs_indexed(dx_stream_t st, int64_t B, const int64_t* idx, const float* Xg, const float* Pg,
                        const uint64_t* adjg, float* Xn, int32_t* cls, uint64_t* adj) {
  foreach (st, B * NN * XP, [=] DX_HD(int64_t i) {
    const int c = (int)(i % XP); const int64_t r = i / XP; const int64_t b = r % B; const int v = (int)(r / B);
    const int64_t g = idx ? idx[b] : b;
    Xn[i] = c < SX ? Xg[(g * NN + v) * SX + c] : 0.f;
  });
  if (cls)
    foreach (st, B * 14, [=] DX_HD(int64_t i) {
      const int64_t b = i % B; const int k = (int)(i / B);
      const int64_t g = idx ? idx[b] : b;
      float val;
      if (k == 0) val = Pg[(g * NN) * NP + 17];
      else if (k == 1) val = Pg[(g * NN) * NP + 18];
      else if (k < 8) val = Pg[(g * NN + (k - 1)) * NP + 19];
      else val = Pg[(g * NN + (k - 7)) * NP + 20];
      cls[i] = (int32_t)val;
    });
  if (adj) foreach (st, B, [=] DX_HD(int64_t b) { adj[b] = adjg[idx ? idx[b] : b]; });
  return check_launch("pack_graphs_indexed");
}

int unpack_graphs(dx_stream_t st, int64_t B, const float* Xn, const float* Pn, float* Xg, float* Pg) {
  foreach (st, B * NN * SX, [=] DX_HD(int64_t i) {
    const int c = (int)(i % SX); const int64_t g = i / SX; const int v = (int)(g % NN); const int64_t b = g / NN;
    Xg[i] = Xn[((int64_t)v * B + b) * XP + c];
  });
  foreach (st, B * NN * NP, [=] DX_HD(int64_t i) {
    const int c = (int)(i % NP); const int64_t g = i / NP; const int v = (int)(g % NN); const int64_t b = g / NN;
    Pg[i] = Pn[((int64_t)v * B + b) * XP + c];
  });
  return check_launch("unpack_graphs");
}

// ---- dxdata.py:174-312 ------------------------------------------------------------------------
DX_HD DX_INLINE int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
DX_HD DX_INLINE float tabf(const uint32_t* t, int i) { union { uint32_t u; float f; } c; c.u = t[i]; return c.f; }

#ifndef DX_EMU
__constant__ uint32_t d_tab32[32];
__constant__ uint32_t d_tab100[100];
#define DXD_TAB32 d_tab32
#define DXD_TAB100 d_tab100
static void upload_data_tables() {
  static unsigned long long done = 0;   // __constant__ memory is per device: one bit per device ordinal
  int dev = 0; cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done & bit) return;
  cudaMemcpyToSymbol(d_tab32, kLogTab32, sizeof(kLogTab32));
  cudaMemcpyToSymbol(d_tab100, kLogTab100, sizeof(kLogTab100));
  done |= bit;
}
#else
#define DXD_TAB32 kLogTab32
#define DXD_TAB100 kLogTab100
static void upload_data_tables() {}
#endif

int voices_to_graphs(dx_stream_t st, int64_t B, const uint8_t* voices, float* Xn, int32_t* cls, uint64_t* adj,
                     float* Xg, float* Pg) {
  upload_data_tables();
  foreach (st, B * NN, [=] DX_HD(int64_t i) {
    const int64_t b = i % B; const int v = (int)(i / B);
    const uint8_t* pz = voices + b * 128;
    float x[XP]; float p[NP];
    for (int c = 0; c < XP; ++c) x[c] = 0.f;
    for (int c = 0; c < NP; ++c) p[c] = 0.f;
    if (v == 0) {                                   // parse_global, dxdata.py:246-300
      for (int c = 0; c < 8; ++c) { const int e = clampi(pz[102 + c], 0, 99); p[c] = (float)e; x[c] = (float)e / 99.f; }
      const int alg = pz[110] % 32;
      const int oks = (pz[111] / 8) % 2, fb = pz[111] % 8;
      const int lfs = clampi(pz[112], 0, 99), lfd = clampi(pz[113], 0, 99), lpmd = clampi(pz[114], 0, 99),
                lamd = clampi(pz[115], 0, 99);
      const int lpms = pz[116] / 16, lfw = clampi((pz[116] / 2) % 8, 0, 5), lks = pz[116] % 2;
      const int tsp = clampi(pz[117], 0, 48);
      p[8] = (float)tsp; p[9] = (float)lfs; p[10] = (float)lfd; p[11] = (float)lpmd; p[12] = (float)lamd;
      p[13] = (float)fb; p[14] = (float)lpms; p[15] = (float)oks; p[16] = (float)lks; p[17] = (float)lfw;
      p[18] = (float)alg;
      x[8] = (float)tsp / 48.f; x[9] = (float)lfs / 99.f; x[10] = (float)lfd / 99.f; x[11] = (float)lpmd / 99.f;
      x[12] = (float)lamd / 99.f; x[13] = (float)fb / 7.f; x[14] = (float)lpms / 7.f; x[15] = (float)oks;
      x[16] = (float)lks; x[17 + lfw] = 1.f;
      if (adj) adj[b] = algo_mask(pz[110]);       // dxdata.py:308 (un-modded key; legal voices are < 32)
      if (cls) { cls[b] = lfw; cls[B + b] = alg; }
    } else {                                        // parse_op, dxdata.py:175-244
      const uint8_t* q = pz + (6 - v) * 17;
      const int lev = clampi(q[14], 0, 99);
      p[0] = (float)lev; x[0] = (float)lev / 99.f;
      for (int c = 0; c < 8; ++c) { const int e = clampi(q[c], 0, 99); p[1 + c] = (float)e; x[1 + c] = (float)e / 99.f; }
      const int bp = clampi(q[8], 0, 99), ld = clampi(q[9], 0, 99), rd = clampi(q[10], 0, 99);
      const int rc = (q[11] / 4) % 4, lc = q[11] % 4;
      const int det = clampi(q[12] / 8, 0, 14), rs = q[12] % 8;
      const int kvs = (q[13] / 4) % 8, ams = q[13] % 4;
      int fc = (q[15] / 2) % 32; const int mode = q[15] % 2;
      const int ff = clampi(q[16], 0, 99);
      if (mode == 0) { x[9] = tabf(DXD_TAB32, fc); x[10] = tabf(DXD_TAB100, ff); }
      else { fc = fc % 4; x[9] = (float)fc / 3.f; x[10] = (float)ff / 99.f; }
      p[9] = (float)fc; p[10] = (float)ff; p[11] = (float)det; p[12] = (float)bp; p[13] = (float)ld; p[14] = (float)rd;
      p[15] = (float)ams; p[16] = (float)kvs; p[17] = (float)rs; p[18] = (float)mode; p[19] = (float)lc;
      p[20] = (float)rc;
      x[11] = (float)det / 14.f; x[12] = (float)bp / 99.f; x[13] = (float)ld / 99.f; x[14] = (float)rd / 99.f;
      x[15] = (float)ams / 3.f; x[16] = (float)kvs / 7.f; x[17] = (float)rs / 7.f; x[18] = (float)mode;
      x[19 + lc] = 1.f; x[23 + rc] = 1.f;
      if (cls) { cls[(int64_t)(2 + v - 1) * B + b] = lc; cls[(int64_t)(8 + v - 1) * B + b] = rc; }
    }
    if (Xn) { float* o = Xn + ((int64_t)v * B + b) * XP; for (int c = 0; c < XP; ++c) o[c] = x[c]; }
    if (Xg) { float* o = Xg + (b * NN + v) * SX; for (int c = 0; c < SX; ++c) o[c] = x[c]; }
    if (Pg) { float* o = Pg + (b * NN + v) * NP; for (int c = 0; c < NP; ++c) o[c] = p[c]; }
  });
  return check_launch("voices_to_graphs");
}

// ---- dxdata.py:341-397: one 128-byte voice per graph -------------------------------------------
int pack_syx(dx_stream_t st, int64_t B, const float* Pg, uint8_t* voices) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    uint8_t* o = voices + b * 128;
    const float* pg = Pg + b * NN * NP;
    int k = 0;
    for (int idx = 6; idx >= 1; --idx) {
      const float* pf = pg + idx * NP;
      int pi[NP];
      for (int c = 0; c < NP; ++c) pi[c] = (int)pf[c];          // .int(): truncation (dxdata.py:348)
      for (int c = 0; c < 8; ++c) o[k++] = (uint8_t)pi[1 + c];
      o[k++] = (uint8_t)pi[12]; o[k++] = (uint8_t)pi[13]; o[k++] = (uint8_t)pi[14];
      o[k++] = (uint8_t)(pi[20] * 4 + pi[19]);
      o[k++] = (uint8_t)(pi[11] * 8 + pi[17]);
      o[k++] = (uint8_t)(pi[16] * 4 + pi[15]);
      o[k++] = (uint8_t)pi[0];
      o[k++] = (uint8_t)(pi[9] * 2 + pi[18]);
      o[k++] = (uint8_t)pi[10];
    }
    int p0[NP];
    for (int c = 0; c < NP; ++c) p0[c] = (int)pg[c];
    for (int c = 0; c < 8; ++c) o[k++] = (uint8_t)p0[c];
    o[k++] = (uint8_t)p0[18];
    o[k++] = (uint8_t)(p0[15] * 8 + p0[13]);
    o[k++] = (uint8_t)p0[9]; o[k++] = (uint8_t)p0[10]; o[k++] = (uint8_t)p0[11]; o[k++] = (uint8_t)p0[12];
    o[k++] = (uint8_t)(p0[14] * 16 + p0[17] * 2 + p0[16]);
    o[k++] = (uint8_t)p0[8];
    const uint8_t name[10] = {68, 88, 45, 86, 65, 69, 46, 46, 46, 46};  // "DX-VAE...."
    for (int c = 0; c < 10; ++c) o[k++] = name[c];
  });
  return check_launch("pack_syx");
}

// ---- torch.optim.AdamW single-tensor update on the flat blob (model.py:375, 386) ------------------
int adamw_step(dx_stream_t st, int64_t n, float* w, const float* g, float* m, float* v, float lr, float b1, float b2,
               float eps, float wd, int64_t step, float gscale) {
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2s = sqrtf(1.f - powf(b2, (float)step));
  const float step_size = lr / bc1;
  foreach (st, n, [=] DX_HD(int64_t i) {
    const float gr = g[i] * gscale;
    float p = w[i];
    p *= (1.f - lr * wd);
    const float mi = m[i] + (gr - m[i]) * (1.f - b1);          // lerp_
    const float vi = v[i] * b2 + (1.f - b2) * gr * gr;          // mul_().addcmul_()
    const float denom = sqrtf(vi) / bc2s + eps;
    p -= step_size * (mi / denom);
    w[i] = p; m[i] = mi; v[i] = vi;
  });
  return check_launch("adamw_step");
}

void pad_wih(dx_stream_t st, const float* W, int K, float* Wp) {
  foreach (st, (int64_t)G3 * XP, [=] DX_HD(int64_t i) {
    const int c = (int)(i % XP); const int64_t n = i / XP;
    Wp[i] = c < K ? W[n * K + c] : 0.f;
  });
}
// dWp rows are in (n, r, z) gate order (the D4 view the cell backward hands to the weight_ih products), the blob
// is (r, z, n): blob row n takes padded row (n + H) mod 3H.
void unpad_add_wih(dx_stream_t st, const float* dWp, int K, float* dW) {
  foreach (st, (int64_t)G3 * K, [=] DX_HD(int64_t i) {
    const int c = (int)(i % K); const int64_t n = i / K;
    dW[i] += dWp[((n + H) % G3) * XP + c];
  });
}
void mask_features(dx_stream_t st, int64_t rows, int B, const int* row_ids, int row_base, const uint64_t* adj,
                   const float* X, float* XL) {
  foreach (st, rows * (XP / 4), [=] DX_HD(int64_t i) {
    const int64_t m = i / (XP / 4); const int c = (int)(i % (XP / 4)) * 4;
    const int64_t r = (row_ids ? row_ids[m] : m) + row_base;
    const int b = (int)(r % B), v = (int)(r / B);
    const float s = (float)abit(adj[b], v, v);
    const float4 x = ld4f(X + m * XP + c);
    st4f(XL + m * XP + c, make_float4(s * x.x, s * x.y, s * x.z, s * x.w));
  });
}

int reparameterize(dx_stream_t st, int64_t n, const float* mu, const float* sd, const float* eps, float* z) {
  foreach (st, n, [=] DX_HD(int64_t i) { z[i] = mu[i] + sd[i] * eps[i]; });
  return check_launch("reparameterize");
}

}  // namespace dx
