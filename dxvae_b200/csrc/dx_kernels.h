// Element-wise kernels of the hot path: GRU cell (fwd/bwd), gated-sum message
// aggregation with feedback-edge handling (fwd/bwd), loss heads, quantisers.
// Every kernel is a per-element functor launched through dx::foreach, so the same
// code runs on the GPU (grid-stride, 128-bit loads, coalesced along the hidden
// dimension) and, in tests/emu builds, serially on the CPU.
//
// Row addressing.  Node-major buffers hold one row per (node v, graph b):
// global row r = v*B + b.  A launch covers M rows; compact index m in [0,M):
//     r = (rows ? rows[m] : m) + row_base ;  b = r % B ;  v = r / B
// Buffers flagged "global" are indexed by r, the others by m.
#pragma once
#include "dx_rt.h"

namespace dx {

DX_HD DX_INLINE int abit(uint64_t a, int s, int d) { return (int)((a >> (s * 7 + d)) & 1ull); }
DX_HD DX_INLINE float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
DX_HD DX_INLINE void st4f(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
DX_HD DX_INLINE float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

struct RowMap {
  int M; int B; const int* rows; int row_base;
  DX_HD DX_INLINE int r(int m) const { return (rows ? rows[m] : m) + row_base; }
};

enum { S_ONE = 0, S_ZERO = 1, S_SELF = 2 };  // x multiplier of the GRU input: 1, 0, or the node's self-loop flag

// ------------------------------------------------------------------------------------
// GRU cell forward (torch.nn.GRUCell, gate order r,z,n — model.py:186,192,193)
//   gx, gh: raw products x W_ih^T and h W_hh^T (no bias), compact [M,1536]; gh==NULL means h=0
//   r = sig((s*gx_r+b_ir)+(gh_r+b_hr)); z likewise; n = tanh((s*gx_n+b_in) + r*(gh_n+b_hn))
//   h' = (1-z)*n + z*h
// gates (optional) receives r,z,n,nh=(gh_n+b_hn) as [.,2048] for the backward pass.
// ------------------------------------------------------------------------------------
struct CellFwd {
  RowMap rm; const float* gx; const float* gh; const float* bih; const float* bhh;
  const float* hprev; int hprev_global; float* hout; int hout_global; float* gates; int gates_global;
  int smode; const uint64_t* adj;
  int gx_by_graph = 0;        // 1: gx is [B,1536] indexed by the row's graph b (= r % B), not by m
  float* hout2 = nullptr;     // optional second copy of h', [B,512] indexed by b (current state of the node)
  float* hout3 = nullptr;     // optional third copy, likewise
  int gh_by_graph = 0;        // 1: gh is [B,1536] indexed by the row's graph b
  int hprev_by_graph = 0;     // 1: hprev is [B,512] indexed by the row's graph b
  int hout23_local = 0;       // 1: hout2 / hout3 are [M,512] indexed by m like hout (all nodes' first propagates in one launch)
};

inline void cell_fwd(dx_stream_t st, const CellFwd& a) {
  foreach (st, (int64_t)a.rm.M * (H / 4), [=] DX_HD(int64_t idx) {
    const int m = (int)(idx / (H / 4)), n = (int)(idx % (H / 4)) * 4;
    const int r = a.rm.r(m);
    float s = 1.f;
    if (a.smode == S_ZERO) s = 0.f;
    else if (a.smode == S_SELF) { const int b = r % a.rm.B, v = r / a.rm.B; s = (float)abit(a.adj[b], v, v); }
    // s == 0 (first propagate of a new node; looper input of a node without a self-loop): s*gx is +-0 whatever
    // gx holds, so the 6 KB/row input product is not read at all
    float4 xr = f4zero(), xz = f4zero(), xn = f4zero();
    if (s != 0.f) {
      const float* gx = a.gx + (int64_t)(a.gx_by_graph ? r % a.rm.B : m) * G3 + n;
      xr = ld4f(gx); xz = ld4f(gx + H); xn = ld4f(gx + 2 * H);
    }
    float4 hr = f4zero(), hz = f4zero(), hn = f4zero(), hp = f4zero();
    if (a.gh) { const float* gh = a.gh + (int64_t)(a.gh_by_graph ? r % a.rm.B : m) * G3 + n; hr = ld4f(gh); hz = ld4f(gh + H); hn = ld4f(gh + 2 * H); }
    if (a.hprev) hp = ld4f(a.hprev + (int64_t)(a.hprev_by_graph ? r % a.rm.B : (a.hprev_global ? r : m)) * H + n);
    const float4 bir = ld4f(a.bih + n), biz = ld4f(a.bih + H + n), bin = ld4f(a.bih + 2 * H + n);
    const float4 bhr = ld4f(a.bhh + n), bhz = ld4f(a.bhh + H + n), bhn = ld4f(a.bhh + 2 * H + n);
    float4 R, Zg, Ng, NH, O;
#define DX_CELL(c)                                                      \
    {                                                                   \
      const float rr = sigmoidf_((s * xr.c + bir.c) + (hr.c + bhr.c));  \
      const float zz = sigmoidf_((s * xz.c + biz.c) + (hz.c + bhz.c));  \
      const float nh = hn.c + bhn.c;                                    \
      const float nn = tanhf((s * xn.c + bin.c) + rr * nh);             \
      R.c = rr; Zg.c = zz; Ng.c = nn; NH.c = nh;                        \
      O.c = (1.f - zz) * nn + zz * hp.c;                                \
    }
    DX_CELL(x) DX_CELL(y) DX_CELL(z) DX_CELL(w)
#undef DX_CELL
    st4f(a.hout + (int64_t)(a.hout_global ? r : m) * H + n, O);
    const int64_t o23 = a.hout23_local ? m : r % a.rm.B;
    if (a.hout2) st4f(a.hout2 + o23 * H + n, O);
    if (a.hout3) st4f(a.hout3 + o23 * H + n, O);
    if (a.gates) {
      float* g = a.gates + (int64_t)(a.gates_global ? r : m) * (4 * H) + n;
      st4f(g, R); st4f(g + H, Zg); st4f(g + 2 * H, Ng); st4f(g + 3 * H, NH);
    }
  });
}

// GRU cell backward.  dh: gradient of h' ; outputs (compact [M,.]):
//   ONE buffer D4[M, 4H] = [da_n | da_r | da_z | da_n * r] serves both gate-gradient matrices as strided views
//   (8 KB/row written instead of 12):
//     dgx = D4[:, 0:3H]   = [da_n, da_r, da_z]      weight_ih gradient; its rows come out in (n, r, z) gate order and
//                                                    unpad_add_wih() folds them back into the (r, z, n) blob order
//     dgh = D4[:, H:4H]   = [da_r, da_z, da_n * r]  bias_hh / weight_hh gradient, and dh_prev += dgh W_hh
//   a.dgx points at D4 (a.dgh == a.dgx + H, row pitch 4H for both; a.dgxs is unused)
//   dhp  = dh * z                      (direct path to h_prev)
struct CellBwd {
  RowMap rm; const float* dh; int dh_global; const float* gates; int gates_global; const float* hprev;
  int hprev_global; float* dgx; float* dgxs; float* dgh; float* dhp; int smode; const uint64_t* adj;
  int dhp_global = 0;
};

// One (row m, 4 hidden units at n) element of the backward pass; returns the bias-gradient
// contributions (ar, az, an, an*r) of that element.
DX_HD DX_INLINE void cell_bwd_elem(const CellBwd& a, int m, int n, float4& ar, float4& az, float4& an, float4& anr) {
  const int r = a.rm.r(m);   // (the gate gradients do not depend on the input multiplier s: it only scales x in the weight_ih product)
  const float4 d = ld4f(a.dh + (int64_t)(a.dh_global ? r : m) * H + n);
  const float* g = a.gates + (int64_t)(a.gates_global ? r : m) * (4 * H) + n;
  const float4 R = ld4f(g), Zg = ld4f(g + H), Ng = ld4f(g + 2 * H), NH = ld4f(g + 3 * H);
  float4 hp = f4zero();
  if (a.hprev) hp = ld4f(a.hprev + (int64_t)(a.hprev_global ? r : m) * H + n);
  float4 dp;
#define DX_CELLB(c)                                            \
  {                                                            \
    const float dn = d.c * (1.f - Zg.c);                       \
    const float dz = d.c * (hp.c - Ng.c);                      \
    const float dan = dn * (1.f - Ng.c * Ng.c);                \
    const float dr = dan * NH.c;                               \
    ar.c = dr * R.c * (1.f - R.c);                             \
    az.c = dz * Zg.c * (1.f - Zg.c);                           \
    an.c = dan; anr.c = dan * R.c; dp.c = d.c * Zg.c;          \
  }
  DX_CELLB(x) DX_CELLB(y) DX_CELLB(z) DX_CELLB(w)
#undef DX_CELLB
  float* o = a.dgx + (int64_t)m * (4 * H) + n;
  st4f(o, an); st4f(o + H, ar); st4f(o + 2 * H, az); st4f(o + 3 * H, anr);
  if (a.dhp) st4f(a.dhp + (int64_t)(a.dhp_global ? r : m) * H + n, dp);
}

#ifndef DX_EMU
// Block = 4 row groups x 128 threads (the 512 hidden units, 4 per thread); blocks stride over rows.
// Every thread keeps the column sums of its 4 units in registers (bias gradients = column sums of
// dgx / dgh), the 4 row groups are folded through shared memory, and 128 threads issue the atomics:
// no second pass over 12 KB/row, and only 2 blocks/SM worth of atomics per address.
static __global__ void __launch_bounds__(512) k_cell_bwd(const CellBwd a, float* __restrict__ dbih, float* __restrict__ dbhh) {
  pdl_wait();
  __shared__ float red[3][16][128];
  const int q = threadIdx.x & 127, rg = threadIdx.x >> 7;
  const int n = q * 4;
  float4 sr = f4zero(), sz = f4zero(), sn = f4zero(), snr = f4zero();
  for (int m = blockIdx.x * 4 + rg; m < a.rm.M; m += gridDim.x * 4) {
    float4 ar, az, an, anr;
    cell_bwd_elem(a, m, n, ar, az, an, anr);
    sr.x += ar.x; sr.y += ar.y; sr.z += ar.z; sr.w += ar.w;
    sz.x += az.x; sz.y += az.y; sz.z += az.z; sz.w += az.w;
    sn.x += an.x; sn.y += an.y; sn.z += an.z; sn.w += an.w;
    snr.x += anr.x; snr.y += anr.y; snr.z += anr.z; snr.w += anr.w;
  }
  if (!dbih) return;
  const float v[16] = {sr.x, sr.y, sr.z, sr.w, sz.x, sz.y, sz.z, sz.w, sn.x, sn.y, sn.z, sn.w, snr.x, snr.y, snr.z, snr.w};
  if (rg > 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e) red[rg - 1][e][q] = v[e];
  }
  __syncthreads();
  if (rg == 0) {
    float t[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) t[e] = v[e] + red[0][e][q] + red[1][e][q] + red[2][e][q];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(dbih + n + e, t[e]);          atomicAdd(dbhh + n + e, t[e]);            // r gate
      atomicAdd(dbih + H + n + e, t[4 + e]);  atomicAdd(dbhh + H + n + e, t[4 + e]);    // z gate
      atomicAdd(dbih + 2 * H + n + e, t[8 + e]);                                         // n gate, input side
      atomicAdd(dbhh + 2 * H + n + e, t[12 + e]);                                        // n gate, hidden side (x r)
    }
  }
}
#endif

// dbih / dbhh (optional): bias_ih / bias_hh gradients, accumulated (+=) with the column sums of dgx / dgh.
inline void cell_bwd(dx_stream_t st, const CellBwd& a, float* dbih = nullptr, float* dbhh = nullptr) {
  if (a.rm.M <= 0) return;
#ifndef DX_EMU
  int blocks = (a.rm.M + 3) / 4;
  if (blocks > 148 * 2) blocks = 148 * 2;
  launch_k(k_cell_bwd, dim3(blocks), dim3(512), 0, st, 1, a, dbih, dbhh);
  ++g_launches;
#else
  for (int m = 0; m < a.rm.M; ++m)
    for (int n = 0; n < H; n += 4) {
      float4 ar, az, an, anr;
      cell_bwd_elem(a, m, n, ar, az, an, anr);
      if (dbih) {
        const float vi[12] = {ar.x, ar.y, ar.z, ar.w, az.x, az.y, az.z, az.w, an.x, an.y, an.z, an.w};
        const float vh[12] = {ar.x, ar.y, ar.z, ar.w, az.x, az.y, az.z, az.w, anr.x, anr.y, anr.z, anr.w};
        for (int g = 0; g < 3; ++g)
          for (int e = 0; e < 4; ++e) { dbih[g * H + n + e] += vi[g * 4 + e]; dbhh[g * H + n + e] += vh[g * 4 + e]; }
      }
    }
  ++g_launches;
#endif
}

// ------------------------------------------------------------------------------------
// Gated-sum neighbour aggregation (model.py:163-181) on pre-projected neighbour states.
//   Pg[x-row, n] = W_g[n,:512] h_x ("in" half), Pg[x-row, 512+n] = W_g[n,512:] h_x ("out" half)
//   Pm likewise for the bias-free mapper.  A half is only read where its edge flag is set, so the
//   producer may skip the half a row never needs (DESIGN.md "projection halves").
//   message x->v:  i=[x->v], o=[v->x] ;  m = sig(i*Pg_in + o*Pg_out + b_g) * (i*Pm_in + o*Pm_out)
//   (i=o=0 gives exactly 0 in the reference because the mapper has no bias: skipped.)
// Forward: target rows (RowMap), neighbours x = x_lo..x_hi in steps of +1 (encode: v+1..6,
// decode step: x_lo=x_hi=vj).  x_lo<0 means "v+1".  accum=1 adds to the existing Hin.
// P buffers are global-row ([7B,1024]); hin is global (hin_global) or compact.
// ------------------------------------------------------------------------------------
struct MsgFwd {
  RowMap rm; const float* Pg; const float* Pm; const float* bg; const uint64_t* adj; float* hin; int hin_global;
  int x_lo, x_hi; int accum;   // accum: see msg_fwd
  const int* pos = nullptr;   // optional: row of neighbour (x,b) in Pg/Pm is pos[x*B+b] instead of x*B+b
  int hin_by_graph = 0;       // 1: hin is [B,512] indexed by the row's graph b
  float* hin_copy = nullptr;  // optional compact copy [M,512] of the updated aggregate
};

inline void msg_fwd(dx_stream_t st, const MsgFwd& a) {
  foreach (st, (int64_t)a.rm.M * (H / 4), [=] DX_HD(int64_t idx) {
    const int m = (int)(idx / (H / 4)), n = (int)(idx % (H / 4)) * 4;
    const int r = a.rm.r(m);
    const int b = r % a.rm.B, v = r / a.rm.B;
    const uint64_t A = a.adj[b];
    float* out = a.hin + (int64_t)(a.hin_by_graph ? b : (a.hin_global ? r : m)) * H + n;
    const int lo = a.x_lo < 0 ? v + 1 : a.x_lo, hi = a.x_lo < 0 ? NN - 1 : a.x_hi;
    // accum: 0 = store, 1 = add to what hin holds, 2 (decoder steps, one neighbour x per call, x descending from v-1) =
    // add only if an earlier call already wrote this row, i.e. the graph has an edge between v and a neighbour in
    // (x, v): the first message of a row then needs no zero-initialised aggregate
    bool add = a.accum == 1;
    if (a.accum == 2)
      for (int xe = hi + 1; xe < v; ++xe) add = add || (abit(A, xe, v) | abit(A, v, xe));
    float4 acc = add ? ld4f(out) : f4zero();
    const float4 bg = ld4f(a.bg + n);
    for (int x = lo; x <= hi; ++x) {
      const float fi = (float)abit(A, x, v), fo = (float)abit(A, v, x);
      if (fi == 0.f && fo == 0.f) continue;
      const int64_t xr = a.pos ? (int64_t)a.pos[x * a.rm.B + b] : (int64_t)x * a.rm.B + b;
      // a half is only read where its flag is set: the other one may not have been computed for this row
      float4 gi = f4zero(), go = f4zero(), pi = f4zero(), po = f4zero();
      if (fi != 0.f) { gi = ld4f(a.Pg + xr * (2 * H) + n); pi = ld4f(a.Pm + xr * (2 * H) + n); }
      if (fo != 0.f) { go = ld4f(a.Pg + xr * (2 * H) + H + n); po = ld4f(a.Pm + xr * (2 * H) + H + n); }
      acc.x += sigmoidf_((fi * gi.x + fo * go.x) + bg.x) * (fi * pi.x + fo * po.x);
      acc.y += sigmoidf_((fi * gi.y + fo * go.y) + bg.y) * (fi * pi.y + fo * po.y);
      acc.z += sigmoidf_((fi * gi.z + fo * go.z) + bg.z) * (fi * pi.z + fo * po.z);
      acc.w += sigmoidf_((fi * gi.w + fo * go.w) + bg.w) * (fi * pi.w + fo * po.w);
    }
    st4f(out, acc);
    if (a.hin_copy) st4f(a.hin_copy + (int64_t)m * H + n, acc);
  });
}

// Backward, from the SOURCE row's perspective:
// source rows (RowMap) x; targets v = v_lo..v_hi (encode: 0..x-1 when v_lo<0).
// dhin row of target v: (v*dh_vstride + b).  Writes/accumulates dPg, dPm ([.,1024]) at the source row (global r or
// compact m).  The gate-bias gradient only matters summed over rows: every thread keeps the column sums of its 4
// hidden units in registers and the block adds them to dbias atomically (no per-row buffer, no second pass).
struct MsgBwd {
  RowMap rm; const float* Pg; const float* Pm; const float* bg; const uint64_t* adj; const float* dhin;
  int64_t dh_vstride; float* dPg; float* dPm; float* dbias; int out_global; int v_lo, v_hi; int accum;
  const int* pos = nullptr;   // optional: dhin row of target (v,b) is pos[v*B+b]
  int p_compact = 0;          // 1: this row's own Pg/Pm are indexed by m (pointer pre-offset), else by r
  // halves nobody reads are not written (DESIGN.md "projection halves"):
  int out_rows = -1;          // >= 0 (encoder, accum = 0): only rows m < out_rows can have an "out" flag; the others skip that half
  int lazy_in = 0;            // 1 (decoder, accum = 1): the "in" half is read-modify-written only when an "in" flag is set
};

// One (row m, 4 hidden units at n) element; returns this call's gate pre-activation gradient (bias contribution).
DX_HD DX_INLINE float4 msg_bwd_elem(const MsgBwd& a, int m, int n) {
  const int r = a.rm.r(m);
  const int b = r % a.rm.B, x = r / a.rm.B;
  const uint64_t A = a.adj[b];
  const int64_t o = a.out_global ? r : m;
  float* dgp = a.dPg + o * (2 * H) + n; float* dpp = a.dPm + o * (2 * H) + n;
  float4 dgi = f4zero(), dgo = f4zero(), dpi = f4zero(), dpo = f4zero(), db = f4zero();
  bool any_in = false;
  if (a.lazy_in) {
    const int l0 = a.v_lo < 0 ? 0 : a.v_lo, h0 = a.v_lo < 0 ? x - 1 : a.v_hi;
    for (int v = l0; v <= h0; ++v) any_in = any_in || abit(A, x, v);
  }
  const bool touch_in = !a.lazy_in || any_in;
  if (a.accum) {
    if (touch_in) { dgi = ld4f(dgp); dpi = ld4f(dpp); }
    dgo = ld4f(dgp + H); dpo = ld4f(dpp + H);
  }
  const int lo = a.v_lo < 0 ? 0 : a.v_lo, hi = a.v_lo < 0 ? x - 1 : a.v_hi;
  const int64_t own = a.p_compact ? m : r;
  const float* gp = a.Pg + own * (2 * H) + n; const float* pp = a.Pm + own * (2 * H) + n;
  const float4 bg = ld4f(a.bg + n);
  for (int v = lo; v <= hi; ++v) {
    const float fi = (float)abit(A, x, v), fo = (float)abit(A, v, x);
    if (fi == 0.f && fo == 0.f) continue;
    const int64_t trow = a.pos ? (int64_t)a.pos[v * a.rm.B + b] : (int64_t)v * a.dh_vstride + b;
    const float4 dh = ld4f(a.dhin + trow * H + n);
    float4 gi = f4zero(), go = f4zero(), pi = f4zero(), po = f4zero();
    if (fi != 0.f) { gi = ld4f(gp); pi = ld4f(pp); }
    if (fo != 0.f) { go = ld4f(gp + H); po = ld4f(pp + H); }
#define DX_MSGB(c)                                                                \
    {                                                                             \
      const float s_ = sigmoidf_((fi * gi.c + fo * go.c) + bg.c), c_ = fi * pi.c + fo * po.c; \
      const float da = dh.c * c_ * s_ * (1.f - s_), dc = dh.c * s_;               \
      dgi.c += fi * da; dgo.c += fo * da; dpi.c += fi * dc; dpo.c += fo * dc; db.c += da; \
    }
    DX_MSGB(x) DX_MSGB(y) DX_MSGB(z) DX_MSGB(w)
#undef DX_MSGB
  }
  if (touch_in) { st4f(dgp, dgi); st4f(dpp, dpi); }
  if (a.out_rows < 0 || m < a.out_rows) { st4f(dgp + H, dgo); st4f(dpp + H, dpo); }
  return db;
}

#ifndef DX_EMU
// Block = 4 row groups x 128 threads (the 512 hidden units, 4 per thread); blocks stride over the rows (as k_cell_bwd).
static __global__ void __launch_bounds__(512, 2) k_msg_bwd(const MsgBwd a) {
  pdl_wait();
  __shared__ float red[3][4][128];
  const int q = threadIdx.x & 127, rg = threadIdx.x >> 7;
  float4 s = f4zero();
  for (int m = blockIdx.x * 4 + rg; m < a.rm.M; m += gridDim.x * 4) {
    const float4 d = msg_bwd_elem(a, m, q * 4);
    s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
  }
  if (!a.dbias) return;
  if (rg > 0) { red[rg - 1][0][q] = s.x; red[rg - 1][1][q] = s.y; red[rg - 1][2][q] = s.z; red[rg - 1][3][q] = s.w; }
  __syncthreads();
  if (rg == 0) {
    atomicAdd(a.dbias + q * 4 + 0, s.x + red[0][0][q] + red[1][0][q] + red[2][0][q]);
    atomicAdd(a.dbias + q * 4 + 1, s.y + red[0][1][q] + red[1][1][q] + red[2][1][q]);
    atomicAdd(a.dbias + q * 4 + 2, s.z + red[0][2][q] + red[1][2][q] + red[2][2][q]);
    atomicAdd(a.dbias + q * 4 + 3, s.w + red[0][3][q] + red[1][3][q] + red[2][3][q]);
  }
}
#endif

inline void msg_bwd(dx_stream_t st, const MsgBwd& a) {
  if (a.rm.M <= 0) return;
#ifndef DX_EMU
  int blocks = (a.rm.M + 3) / 4;
  if (blocks > 148 * 2) blocks = 148 * 2;
  launch_k(k_msg_bwd, dim3(blocks), dim3(512), 0, st, 1, a);
  ++g_launches;
#else
  for (int m = 0; m < a.rm.M; ++m)
    for (int n = 0; n < H; n += 4) {
      const float4 d = msg_bwd_elem(a, m, n);
      if (a.dbias) { a.dbias[n] += d.x; a.dbias[n + 1] += d.y; a.dbias[n + 2] += d.z; a.dbias[n + 3] += d.w; }
    }
  ++g_launches;
#endif
}

// ------------------------------------------------------------------------------------
// small generic element-wise helpers
// ------------------------------------------------------------------------------------
// Row gather / scatter helpers for the compacted teacher-forced steps (rows[m] = graph index).
//   gather_rows     : dst[m,:] = src[rows[m],:]            ; zero_src also clears the source rows
//   scatter_rows    : dst[rows[m],:] = src[m,:]            ; add=1 accumulates
inline void gather_rows(dx_stream_t st, int M, int C, const int* rows, float* src, float* dst, int zero_src, int ld_src = 0) {
  if (ld_src == 0) ld_src = C;                       // row pitch of src (a column block of a wider matrix when > C)
  foreach (st, (int64_t)M * (C / 4), [=] DX_HD(int64_t idx) {
    const int m = (int)(idx / (C / 4)), c = (int)(idx % (C / 4)) * 4;
    float* sp = src + (int64_t)rows[m] * ld_src + c;
    st4f(dst + (int64_t)m * C + c, ld4f(sp));
    if (zero_src) st4f(sp, f4zero());
  });
}
inline void scatter_rows(dx_stream_t st, int M, int C, const int* rows, const float* src, float* dst, int add, int ld_dst = 0) {
  if (ld_dst == 0) ld_dst = C;
  foreach (st, (int64_t)M * (C / 4), [=] DX_HD(int64_t idx) {
    const int m = (int)(idx / (C / 4)), c = (int)(idx % (C / 4)) * 4;
    float* dp = dst + (int64_t)rows[m] * ld_dst + c;
    float4 v = ld4f(src + (int64_t)m * C + c);
    if (add) { const float4 o = ld4f(dp); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    st4f(dp, v);
  });
}
// e[m,j] = relu(u[m,j] + q[m,j])
inline void add_relu(dx_stream_t st, int64_t n4, const float* u, const float* q, float* e) {
  foreach (st, n4, [=] DX_HD(int64_t i) {
    const float4 a = ld4f(u + 4 * i), b = ld4f(q + 4 * i);
    st4f(e + 4 * i, make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f)));
  });
}
// y[i] (+)= x[i]
inline void add_inplace(dx_stream_t st, int64_t n4, float* y, const float* x) {
  foreach (st, n4, [=] DX_HD(int64_t i) {
    float4 a = ld4f(y + 4 * i); const float4 b = ld4f(x + 4 * i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; st4f(y + 4 * i, a);
  });
}
// dpre[m,j] = (act[m,j] > 0) * sum_c dl[m,c] * W2[c,j]   (backward of  l = relu-act . W2^T, C = 1 or 2 outputs)
// optionally also acc[m,j] += dpre[m,j]
inline void relu_head_bwd(dx_stream_t st, int M, int N, int C, const float* act, const float* dl, int lddl,
                          const float* W2, float* dpre, float* acc, float* acc2 = nullptr) {
  foreach (st, (int64_t)M * (N / 4), [=] DX_HD(int64_t idx) {
    const int m = (int)(idx / (N / 4)), j = (int)(idx % (N / 4)) * 4;
    const float4 a = ld4f(act + (int64_t)m * N + j);
    float4 g = f4zero();
    for (int c = 0; c < C; ++c) {
      const float d = dl[(int64_t)m * lddl + c];
      const float4 w = ld4f(W2 + (int64_t)c * N + j);
      g.x += d * w.x; g.y += d * w.y; g.z += d * w.z; g.w += d * w.w;
    }
    g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f; g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
    if (dpre) st4f(dpre + (int64_t)m * N + j, g);
    if (acc2) { float4 q = ld4f(acc2 + (int64_t)m * N + j); q.x += g.x; q.y += g.y; q.z += g.z; q.w += g.w;
                st4f(acc2 + (int64_t)m * N + j, q); }
    if (acc) { float4 q = ld4f(acc + (int64_t)m * N + j); q.x += g.x; q.y += g.y; q.z += g.z; q.w += g.w;
               st4f(acc + (int64_t)m * N + j, q); }
  });
}
// d[m,j] *= (act[m,j] > 0)
inline void relu_mask(dx_stream_t st, int64_t n4, float* d, const float* act) {
  foreach (st, n4, [=] DX_HD(int64_t i) {
    float4 g = ld4f(d + 4 * i); const float4 a = ld4f(act + 4 * i);
    g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f; g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
    st4f(d + 4 * i, g);
  });
}
// d[i] *= (1 - y[i]^2)      (tanh backward, y = tanh output)
inline void tanh_bwd(dx_stream_t st, int64_t n, float* d, const float* y) {
  foreach (st, n, [=] DX_HD(int64_t i) { d[i] *= (1.f - y[i] * y[i]); });
}
// d[i] *= sech^2(x[i])       (tanh backward from the PRE-activation).  1 - y^2 loses its leading digits when the unit
// saturates (y = 0.9999 stored in fp32 leaves 1 - y^2 good to 3e-4 relative); sech^2 = 4a / (1 + a)^2, a = exp(-2|x|),
// keeps full relative precision.  Used for z_to_h, whose units saturate in a trained model (|mu| reaches 17) and whose
// bias gradient is a heavily cancelling sum over the batch.
inline void tanh_bwd_pre(dx_stream_t st, int64_t n, float* d, const float* x) {
  foreach (st, n, [=] DX_HD(int64_t i) {
    const float a = expf(-2.f * fabsf(x[i])), q = 1.f + a;
    d[i] *= 4.f * a / (q * q);
  });
}

}  // namespace dx
