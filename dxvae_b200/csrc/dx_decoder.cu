// Autoregressive decoder: greedy generation (model.py:214-253, quantisers :87-149) and the
// teacher-forced ELBO of model.py:270-367 with its hand-written backward (model.py:385).
//
// Both walk the same 34-propagate schedule (root, then per operator vi: P1, self-loop
// decision, P2, and one re-propagate per lower node vj).  Exact re-arrangements used here
// (DESIGN.md "decoder schedule"):
//   * x W_ih^T of node vi is computed once and reused by all (2+vi) propagates;
//   * P1/P2 have H_in = 0, so the combiner needs no hidden product and both share the
//     looper's hidden product;
//   * a finished node's gate/mapper projections (Pg,Pm) and its half of the edge-MLP first
//     layer (Q = h_j W_e0[:,512:]^T + b) are computed once and reused by every later node;
//   * the aggregated input of node vi is kept as a running sum, one message added per step
//     (same left-to-right order as the reference's slot sum).
#include "dx_engine.h"
#include "dx_tables.h"

namespace dx {

// step_ptr (HOST, NLIST+1; training with the compacted step schedule): the per-step buffers hold one row per ACTIVE graph
// of the step, so they are sized by the schedule's row counts instead of B (about 30 % of the (graph, step) pairs on
// dataset-like topologies), the self-loop propagate's gates by the self-loop rows, and a step's E1 only holds the
// 256-byte relu bit-mask of the fused edge head.  Without it (dense replay, greedy decode) every buffer has B rows.
// Teacher forcing never feeds a parameter head's output back (x_vi is the true X), so the heads of nodes 1..6 can run as
// ONE 3-layer pass over the 6B finished node states after the node loop, and their backward as one pass before it
// (18 + 54 product launches become 3 + 9).  It needs the two relu-gradient scratch matrices at 6B rows instead of B
// (+40 KB per patch), so it is used for batches up to DX_HEADS_BATCH_MAX graphs (default 4096: the small-batch regime,
// where the step is bound by the number of dependent launches); larger batches keep one pass per node.
static bool heads_batched(int64_t B, bool train) { return train && B <= small_batch_max(); }   // (dx_gemm.h)
// Likewise the FIRST propagates of nodes 1..6 (model.py:234-240 / 320-337: no edges yet, H_in = 0): under teacher forcing
// they depend on the true features of the node only, not on the nodes before it, so in the same small-batch regime the
// input products, the combiner / looper cells, the self-loop head and their backward run once over 6B rows (the per-node
// buffers are carved back to back) instead of once per node; only the second propagate (self-loop rows) and the loss of
// the self-loop head stay in the node loop.  Compacted schedules only (the dense replay keeps the per-node form).
static bool p1_batched(int64_t B, bool train, const int32_t* step_ptr) {
  static const bool off = getenv("DX_NO_P1_BATCH") != nullptr;
  return !off && heads_batched(B, train) && step_ptr != nullptr;
}

DecWs carve_dec(Arena& ar, int64_t B, bool train, const int32_t* step_ptr) {
  DecWs w{};
  const size_t b = (size_t)B;
  const bool compact = train && step_ptr != nullptr;
  auto list_rows = [&](int list) -> size_t {
    if (!compact) return b;
    const size_t n = (size_t)(step_ptr[list + 1] - step_ptr[list]);
    return n > 0 ? n : 1;
  };
  // The per-step state / gate buffers of the compacted schedule take EXACTLY their active rows (an empty step takes
  // none; rows are 2 KB / 8 KB, so no alignment padding either): the 21 steps' buffers are then one contiguous matrix
  // in step order, which is what lets the backward pass form each cell's weight gradients as ONE product over all
  // active (graph, step) pairs instead of one per step (decode_bwd_impl).
  auto step_rows_exact = [&](int t) -> size_t { return compact ? (size_t)(step_ptr[t + 1] - step_ptr[t]) : b; };
  w.z = ar.take<float>(b * Z); w.Hinit = ar.take<float>(b * H); w.Hd = ar.take<float>(7 * b * H);
  w.Pg = ar.take<float>(6 * b * 2 * H); w.Pm = ar.take<float>(6 * b * 2 * H); w.Q = ar.take<float>(6 * b * 4 * H);
  w.gh = ar.take<float>(b * G3); w.ghl0 = ar.take<float>((p1_batched(B, train, step_ptr) ? 6 * b : b) * G3); w.Hrun = ar.take<float>(b * H);
  for (int k = 0; k < 3; ++k) w.WihP[k] = ar.take<float>((size_t)G3 * XP);
  if (train) {
    w.U = ar.take<float>(b * 4 * H); w.UC = ar.take<float>(b * 4 * H); w.dHiC = ar.take<float>(b * H);
    for (int k = 0; k < 3; ++k) w.dWihP[k] = ar.take<float>((size_t)G3 * XP);
    w.XL = ar.take<float>(7 * b * XP); w.xc = ar.take<float>(b * XP);
    const size_t nact = compact ? (size_t)(step_ptr[NSTEP] - step_ptr[0]) : 0;
    w.xlS = ar.take<float>((nact ? nact : 1) * XP); w.xiS = ar.take<float>((nact ? nact : 1) * XP);
  }
  auto per_node = [&](float** arr, size_t cols, bool need) {
    float* shared = need ? nullptr : ar.take<float>(b * cols);
    for (int v = 0; v < 7; ++v) arr[v] = need ? ar.take<float>(b * cols) : shared;
  };
  auto per_step = [&](float** arr, size_t cols, bool need, bool by_rows = false) {
    float* shared = need ? nullptr : ar.take<float>(b * cols);
    for (int t = 0; t < NSTEP; ++t) arr[t] = need ? ar.take<float>((by_rows ? step_rows_exact(t) : b) * cols) : shared;
  };
  per_node(w.A1, 2 * H, train); per_node(w.A2, 2 * H, train); per_node(w.L, LD_L, train);
  per_node(w.gxc, G3, train); per_node(w.gxl, G3, train); per_node(w.Hc0, H, train);
  per_node(w.Hi_p1, H, train); per_node(w.Hi_p2, H, train); per_node(w.ES1, 2 * H, train); per_node(w.ls, LD_E, train);
  per_step(w.E1, compact ? 64 : 4 * H, train); per_step(w.l2, LD_E, train);
  per_step(w.Hin, H, train, true); per_step(w.Hc, H, train, true); per_step(w.Hi, H, train, true);
  if (train) {
    w.g_root = ar.take<float>(b * 4 * H);
    per_node(w.dL, LD_L, true); per_node(w.g_c0, 4 * H, true); per_node(w.g_p1, 4 * H, true);
    for (int v = 0; v < 7; ++v)   // second propagate: self-loop rows only when compacted (list NSTEP + v - 1)
      w.g_p2[v] = ar.take<float>((v == 0 ? (size_t)1 : (compact ? list_rows(NSTEP + v - 1) : b)) * 4 * H);
    per_node(w.dls, LD_E, true);
    per_step(w.dl2, LD_E, true); per_step(w.g_c, 4 * H, true, true); per_step(w.g_l, 4 * H, true, true);
    w.rowloss = ar.take<float>(4 * b);
    w.dHd = ar.take<float>(7 * b * H); w.dPg = ar.take<float>(6 * b * 2 * H); w.dPm = ar.take<float>(6 * b * 2 * H);
    w.dQ = ar.take<float>(6 * b * 4 * H); w.dgb = nullptr;   // (gate-bias gradient: summed inside msg_bwd)
    w.dHi = ar.take<float>(b * H); w.dHc = ar.take<float>(b * H); w.dHin = ar.take<float>(b * H);
    w.dHrun = ar.take<float>(b * H); w.dHc0 = ar.take<float>(b * H);
    w.dgx = ar.take<float>(b * 4 * H); w.dgh = w.dgx + H; w.dgxs = nullptr;   // one D4 buffer, two views (CellBwd)
    w.dE1 = ar.take<float>(b * 4 * H); { const size_t hb = heads_batched(B, train) ? 6 * b : b; w.dA1 = ar.take<float>(hb * 2 * H); w.dA2 = ar.take<float>(hb * 2 * H); }
    w.dES1 = ar.take<float>(b * 2 * H); w.dHinit = ar.take<float>(b * H); w.dz = ar.take<float>(b * Z);
    if (p1_batched(B, train, step_ptr)) { w.dHc06 = ar.take<float>(6 * b * H); w.dir6 = ar.take<float>(6 * b * H); w.dES16 = ar.take<float>(6 * b * 2 * H);
      const size_t na = (size_t)(step_ptr[NSTEP] - step_ptr[0]); w.UCS = ar.take<float>((na ? na : 1) * 4 * H);
      w.dU6 = ar.take<float>(6 * b * 4 * H); w.U6 = ar.take<float>(6 * b * 4 * H); }
  } else {
    w.U = ar.take<float>(b * 4 * H); w.UC = ar.take<float>(b * 4 * H);
    w.act_rows = ar.take<int>(b); w.act_cnt = ar.take<int>(64); w.act_flag = ar.take<uint8_t>(b);
    w.Xd = ar.take<float>(7 * b * XP); w.Pn = ar.take<float>(7 * b * XP);
  }
  return w;
}

// ---- numerically stable pieces shared by the loss heads ---------------------------------
DX_HD DX_INLINE float bce_logits(float x, float t) {  // BCEWithLogitsLoss, reduction='none'
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}

// ---- loss heads (thread per graph): value into rowloss[k][b] (+=), gradient of the logits ----
// node 0: model.py:303-308
static void loss_x0(dx_stream_t st, int B, const float* L0, const float* X0, const int32_t* cls, LossW lw,
                    float* rowloss, float* dL) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    const float* l = L0 + b * LD_L; const float* x = X0 + b * XP; float* d = dL + b * LD_L;
    const float ib = lw.inv_batch;
    float acc = 0.f;
    for (int c = 0; c < 15; ++c) {
      const float w = c < 8 ? lw.w_env : (c == 8 ? lw.w_frq : 1.f);
      const float df = l[c] * w - x[c] * w;
      acc += df * df; d[c] = 2.f * df * w * ib;
    }
    for (int c = 15; c < 17; ++c) { acc += bce_logits(l[c], x[c]); d[c] = (sigmoidf_(l[c]) - x[c]) * ib; }
    for (int seg = 0; seg < 2; ++seg) {
      const int lo = seg == 0 ? 17 : 23, n = seg == 0 ? 6 : 32;
      const int tgt = cls[(int64_t)seg * B + b];
      float mx = l[lo];
      for (int c = 1; c < n; ++c) mx = fmaxf(mx, l[lo + c]);
      float se = 0.f;
      for (int c = 0; c < n; ++c) se += expf(l[lo + c] - mx);
      const float lse = mx + logf(se);
      acc += lse - l[lo + tgt];
      for (int c = 0; c < n; ++c) d[lo + c] = (expf(l[lo + c] - lse) - (c == tgt ? 1.f : 0.f)) * ib;
    }
    for (int c = 55; c < LD_L; ++c) d[c] = 0.f;
    rowloss[b] += acc * ib;  // slot 0: loss_X0
  });
}
// operator vi: model.py:323-328
static void loss_xi(dx_stream_t st, int B, int vi, const float* Li, const float* Xi, const int32_t* cls, LossW lw,
                    float* rowloss, float* dL) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    const float* l = Li + b * LD_L; const float* x = Xi + b * XP; float* d = dL + b * LD_L;
    const float ib = lw.inv_batch;
    float acc = 0.f;
    for (int c = 0; c < 18; ++c) {
      const float w = c < 9 ? lw.w_env : (c == 9 ? lw.w_frq : 1.f);
      const float df = l[c] * w - x[c] * w;
      acc += df * df; d[c] = 2.f * df * w * ib;
    }
    acc += bce_logits(l[18], x[18]); d[18] = (sigmoidf_(l[18]) - x[18]) * ib;
    for (int seg = 0; seg < 2; ++seg) {
      const int lo = 19 + 4 * seg;
      const int tgt = cls[(int64_t)(2 + 6 * seg + (vi - 1)) * B + b];
      float mx = l[lo];
      for (int c = 1; c < 4; ++c) mx = fmaxf(mx, l[lo + c]);
      float se = 0.f;
      for (int c = 0; c < 4; ++c) se += expf(l[lo + c] - mx);
      const float lse = mx + logf(se);
      acc += lse - l[lo + tgt];
      for (int c = 0; c < 4; ++c) d[lo + c] = (expf(l[lo + c] - lse) - (c == tgt ? 1.f : 0.f)) * ib;
    }
    for (int c = 27; c < LD_L; ++c) d[c] = 0.f;
    rowloss[(int64_t)B + b] += acc * ib;  // slot 1: loss_Xi
  });
}
// the same for nodes 1..6 in one launch (batched parameter heads): thread per graph walks its six operators in node
// order, so the row's loss_Xi sum is formed exactly as six consecutive loss_xi launches form it
static void loss_xi_all(dx_stream_t st, int B, const float* L1, const float* Xn, const int32_t* cls, LossW lw,
                        float* rowloss, float* dL1) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    const float ib = lw.inv_batch;
    float tot = rowloss[(int64_t)B + b];
    for (int vi = 1; vi < NN; ++vi) {
      const float* l = L1 + ((int64_t)(vi - 1) * B + b) * LD_L; const float* x = Xn + ((int64_t)vi * B + b) * XP;
      float* d = dL1 + ((int64_t)(vi - 1) * B + b) * LD_L;
      float acc = 0.f;
      for (int c = 0; c < 18; ++c) {
        const float w = c < 9 ? lw.w_env : (c == 9 ? lw.w_frq : 1.f);
        const float df = l[c] * w - x[c] * w;
        acc += df * df; d[c] = 2.f * df * w * ib;
      }
      acc += bce_logits(l[18], x[18]); d[18] = (sigmoidf_(l[18]) - x[18]) * ib;
      for (int seg = 0; seg < 2; ++seg) {
        const int lo = 19 + 4 * seg;
        const int tgt = cls[(int64_t)(2 + 6 * seg + (vi - 1)) * B + b];
        float mx = l[lo];
        for (int c = 1; c < 4; ++c) mx = fmaxf(mx, l[lo + c]);
        float se = 0.f;
        for (int c = 0; c < 4; ++c) se += expf(l[lo + c] - mx);
        const float lse = mx + logf(se);
        acc += lse - l[lo + tgt];
        for (int c = 0; c < 4; ++c) d[lo + c] = (expf(l[lo + c] - lse) - (c == tgt ? 1.f : 0.f)) * ib;
      }
      for (int c = 27; c < LD_L; ++c) d[c] = 0.f;
      tot += acc * ib;
    }
    rowloss[(int64_t)B + b] = tot;  // slot 1: loss_Xi
  });
}
// edge heads: model.py:339 (self loop, n=1: target A[vi,vi]) and :363 (n=2: A[vj,vi], A[vi,vj])
static void loss_edge(dx_stream_t st, int B, int vi, int vj, const float* lg, int n, const uint64_t* adj, LossW lw,
                      float* rowloss, float* dlg) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    const uint64_t A = adj[b];
    float acc = 0.f;
    for (int c = 0; c < n; ++c) {
      const float t = (n == 1) ? (float)abit(A, vi, vi) : (c == 0 ? (float)abit(A, vj, vi) : (float)abit(A, vi, vj));
      const float x = lg[b * LD_E + c];
      acc += bce_logits(x, t);
      dlg[b * LD_E + c] = (sigmoidf_(x) - t) * lw.inv_batch;
    }
    rowloss[(int64_t)2 * B + b] += acc * lw.inv_batch;  // slot 2: loss_E
  });
}
// KL(N(0,1) || N(mu,std)) per graph (model.py:365) and the latent gradients.
//   dmu = dz + w*mu/std^2/B ; dstd = dz*eps + w*(1/std - (1+mu^2)/std^3)/B        (dz may be NULL)
void kld_rows(dx_stream_t st, int B, const float* mu, const float* sd, LossW lw, float* rowloss) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    float acc = 0.f;
    for (int k = 0; k < Z; ++k) {
      const float m = mu[b * Z + k], s = sd[b * Z + k];
      const float vr = (1.f / s) * (1.f / s), t1 = (m / s) * (m / s);
      acc += 0.5f * (vr + t1 - 1.f - logf(vr));
    }
    rowloss[(int64_t)3 * B + b] = acc * lw.inv_batch * lw.w_kld;  // slot 3: kld * w_kld
  });
}
void latent_bwd(dx_stream_t st, int B, const float* mu, const float* sd, const float* eps, const float* dz,
                       LossW lw, float* dmu, float* dsd) {
  foreach (st, (int64_t)B * Z, [=] DX_HD(int64_t i) {
    const float m = mu[i], s = sd[i], g = dz[i];
    const float k = lw.w_kld * lw.inv_batch;
    dmu[i] = g + k * m / (s * s);
    dsd[i] = g * eps[i] + k * (1.f / s - (1.f + m * m) / (s * s * s));
  });
}
// loss5 = (sum, x0, xi, e, kld_w): deterministic two-level sum over the B rows of each slot.
#ifndef DX_EMU
__global__ void __launch_bounds__(1024) k_loss_reduce(int B, const float* __restrict__ rowloss, float* __restrict__ out) {
  pdl_wait();
  __shared__ double red[32];
  __shared__ double tot[4];
  for (int k = 0; k < 4; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) s += (double)rowloss[(int64_t)k * B + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      double v = red[threadIdx.x];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) tot[k] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = (float)(tot[0] + tot[1] + tot[2] + tot[3]);
    out[1] = (float)tot[0]; out[2] = (float)tot[1]; out[3] = (float)tot[2]; out[4] = (float)tot[3];
  }
}
void loss_reduce(dx_stream_t st, int B, const float* rowloss, float* out) {
  launch_k(k_loss_reduce, dim3(1), dim3(1024), 0, st, 1, B, rowloss, out);
  ++g_launches;
}
#else
void loss_reduce(dx_stream_t, int B, const float* rowloss, float* out) {
  double t[4];
  for (int k = 0; k < 4; ++k) { t[k] = 0; for (int i = 0; i < B; ++i) t[k] += rowloss[(int64_t)k * B + i]; }
  out[0] = (float)(t[0] + t[1] + t[2] + t[3]);
  for (int k = 0; k < 4; ++k) out[1 + k] = (float)t[k];
}
#endif

// ---- quantisers (greedy decode), thread per graph -----------------------------------------
DX_HD DX_INLINE float tab_f(const uint32_t* t, int i) {
  union { uint32_t u; float f; } c; c.u = t[i]; return c.f;
}
#ifndef DX_EMU
__constant__ uint32_t c_tab32[32];
__constant__ uint32_t c_tab100[100];
#define DX_TAB32 c_tab32
#define DX_TAB100 c_tab100
static void upload_tables() {
  static unsigned long long done = 0;   // __constant__ memory is per device: one bit per device ordinal
  int dev = 0; cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done & bit) return;
  cudaMemcpyToSymbol(c_tab32, kLogTab32, sizeof(kLogTab32));
  cudaMemcpyToSymbol(c_tab100, kLogTab100, sizeof(kLogTab100));
  done |= bit;
}
#else
#define DX_TAB32 kLogTab32
#define DX_TAB100 kLogTab100
static void upload_tables() {}
#endif

// Every quantiser also reports its decision MARGIN in units of the logit x: how far x is from the nearest point where
// the discrete result would change (a rounding tie, a sigmoid at 0.5, an arg-max tie).  The per-graph minimum goes out
// with the decode so that parity checks can tell a rounding tie from a bug (*mg = min(*mg, margin)).
DX_HD DX_INLINE void q_lin(float x, float scale, float* xo, float* po, float* mg) {   // model.py:87-91
  const float t = x * scale;
  float p = rintf(t);
  *mg = fminf(*mg, (0.5f - fabsf(t - p)) / scale);
  p = fminf(fmaxf(p, 0.f), scale);
  *po = p; *xo = p / scale;
}
DX_HD DX_INLINE float q_round_log(float x, float scale, float* mg) {       // model.py:93-96
  const float ls = logf(scale + 1.f);
  const float t = expf(x * ls) - 1.f;
  float p = rintf(t);
  *mg = fminf(*mg, (0.5f - fabsf(t - p)) / (ls * (t + 1.f)));             // dt/dx = ls * (t + 1)
  return fminf(fmaxf(p, 0.f), scale);
}
DX_HD DX_INLINE float q_bool(float x, float* mg) {                          // model.py:100-102: round(sigmoid(x))
  *mg = fminf(*mg, fabsf(x));
  return rintf(sigmoidf_(x));
}
DX_HD DX_INLINE int argmax_n(const float* l, int n, float* mg = nullptr) {
  int best = 0; float bv = l[0], second = -3.0e38f;
  for (int c = 1; c < n; ++c) {
    if (l[c] > bv) { second = bv; bv = l[c]; best = c; }
    else if (l[c] > second) second = l[c];
  }
  if (mg) *mg = fminf(*mg, bv - second);
  return best;
}
// _reg_x0, model.py:109-125.  Writes node-major X row (32 wide, zero padded) and params row (32 wide).
static void reg_x0(dx_stream_t st, int B, const float* L0, float* Xd, float* Pn, float* margins) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    const float* l = L0 + b * LD_L; float* x = Xd + b * XP; float* p = Pn + b * XP;
    float mg = 3.0e38f;
    for (int c = 0; c < XP; ++c) { x[c] = 0.f; p[c] = 0.f; }
    for (int c = 0; c < 15; ++c) {
      const float sc = c == 8 ? 48.f : (c >= 13 ? 7.f : 99.f);
      q_lin(l[c], sc, &x[c], &p[c], &mg);
    }
    for (int c = 15; c < 17; ++c) { const float v = q_bool(l[c], &mg); x[c] = v; p[c] = v; }
    const int lfw = argmax_n(l + 17, 6, &mg);
    x[17 + lfw] = 1.f; p[17] = (float)lfw;
    p[18] = (float)argmax_n(l + 23, 32, &mg);
    if (margins) margins[2 * b + 1] = fminf(margins[2 * b + 1], mg);
  });
}
// _reg_xi, model.py:127-149 (incl. the 23:26 argmax quirk and the per-mode fc/ff quantiser)
static void reg_xi(dx_stream_t st, int B, const float* Li, float* Xd, float* Pn, float* margins) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    const float* l = Li + b * LD_L; float* x = Xd + b * XP; float* p = Pn + b * XP;
    float mg = 3.0e38f;
    for (int c = 0; c < XP; ++c) { x[c] = 0.f; p[c] = 0.f; }
    for (int c = 0; c < 9; ++c) q_lin(l[c], 99.f, &x[c], &p[c], &mg);
    q_lin(l[11], 14.f, &x[11], &p[11], &mg);
    for (int c = 12; c < 15; ++c) q_lin(l[c], 99.f, &x[c], &p[c], &mg);
    q_lin(l[15], 3.f, &x[15], &p[15], &mg);
    q_lin(l[16], 7.f, &x[16], &p[16], &mg); q_lin(l[17], 7.f, &x[17], &p[17], &mg);
    const float mode = q_bool(l[18], &mg);
    x[18] = mode; p[18] = mode;
    const int lc = argmax_n(l + 19, 4, &mg); x[19 + lc] = 1.f; p[19] = (float)lc;
    const int rc = argmax_n(l + 23, 3, &mg); x[23 + rc] = 1.f; p[20] = (float)rc;
    if (mode == 0.f) {
      const float pc = q_round_log(l[9], 31.f, &mg), pf = q_round_log(l[10], 99.f, &mg);
      p[9] = pc; x[9] = tab_f(DX_TAB32, (int)pc);
      p[10] = pf; x[10] = tab_f(DX_TAB100, (int)pf);
    } else {
      q_lin(l[9], 3.f, &x[9], &p[9], &mg); q_lin(l[10], 99.f, &x[10], &p[10], &mg);
    }
    if (margins) margins[2 * b + 1] = fminf(margins[2 * b + 1], mg);
  });
}
// edge decisions (model.py:236-239, 245-250): sigmoid(logit) > 0.5 ; sets adjacency bits, tracks margins
static void decide_edges(dx_stream_t st, int B, int vi, int vj, const float* lg, int n, uint64_t* adj, float* margins) {
  foreach (st, B, [=] DX_HD(int64_t b) {
    uint64_t A = adj[b];
    float mg = margins ? margins[2 * b] : 0.f;
    for (int c = 0; c < n; ++c) {
      const float x = lg[b * LD_E + c];
      const bool on = sigmoidf_(x) > 0.5f;
      int s, d;
      if (n == 1) { s = vi; d = vi; } else if (c == 0) { s = vj; d = vi; } else { s = vi; d = vj; }
      if (on) A |= (1ull << (s * 7 + d));
      mg = fminf(mg, fabsf(x));
    }
    adj[b] = A;
    if (margins) margins[2 * b] = mg;
  });
}

// =============================================================================================
// Fused edge head of a compacted teacher-forced step (model.py:348-351 + its loss :363 + the
// gradients that are local to the head):  e = relu(U + Q_vj) ;  l = e W2^T + b2 ;  BCE against the
// true (vj->vi, vi->vj) flags ;  dl = (sigmoid(l) - t) / B ;  dW2 += dl^T e ;  db2 += sum dl.
// One pass over U and Q (16 KB/row); only the relu bit-mask (256 B/row) and dl are kept for the
// backward pass instead of the 8 KB/row activation.
// =============================================================================================
struct EdgeHeadP {
  int B, vi, vj; const float* U; const float* Q; const float* W2; const float* b2; const uint64_t* adj; float inv_batch;
  float* l2; float* dl2; float* rowloss; uint8_t* mask; float* dW2; float* db2;
  float* db0 = nullptr;   // optional (with dW2): gradient slot of h_to_edge.0.bias; receives sum_b mask * (dl0 W2[0] + dl1 W2[1])
  // greedy generation (model.py:245-250) instead of the teacher-forced loss: decide both edges from the logits
  // (sigmoid > 0.5), set them in adj_out, track the decision margin and flag the graphs that gained an edge
  uint64_t* adj_out = nullptr; float* margins = nullptr; uint8_t* active = nullptr;
};
// greedy finish of one graph: returns nothing; l0 = logit of vj -> vi, l1 = logit of vi -> vj
DX_HD DX_INLINE void edge_head_row_decide(const EdgeHeadP& a, int b, float l0, float l1) {
  a.l2[(int64_t)b * LD_E] = l0; a.l2[(int64_t)b * LD_E + 1] = l1;
  const bool on0 = sigmoidf_(l0) > 0.5f, on1 = sigmoidf_(l1) > 0.5f;
  uint64_t A = a.adj_out[b];
  if (on0) A |= 1ull << (a.vj * 7 + a.vi);
  if (on1) A |= 1ull << (a.vi * 7 + a.vj);
  a.adj_out[b] = A;
  if (a.margins) a.margins[2 * b] = fminf(a.margins[2 * b], fminf(fabsf(l0), fabsf(l1)));
  a.active[b] = (uint8_t)((on0 || on1) ? 1 : 0);
}

DX_HD DX_INLINE void edge_head_row_finish(const EdgeHeadP& a, int b, float l0, float l1, float* dl) {
  const uint64_t A = a.adj[b];
  const float t0 = (float)abit(A, a.vj, a.vi), t1 = (float)abit(A, a.vi, a.vj);
  l0 += a.b2[0]; l1 += a.b2[1];
  a.l2[(int64_t)b * LD_E] = l0; a.l2[(int64_t)b * LD_E + 1] = l1;
  dl[0] = (sigmoidf_(l0) - t0) * a.inv_batch; dl[1] = (sigmoidf_(l1) - t1) * a.inv_batch;
  a.dl2[(int64_t)b * LD_E] = dl[0]; a.dl2[(int64_t)b * LD_E + 1] = dl[1];
  a.rowloss[(int64_t)2 * a.B + b] += (bce_logits(l0, t0) + bce_logits(l1, t1)) * a.inv_batch;
}

#ifndef DX_EMU
constexpr int EH_R = 4;   // rows per block iteration
// 256 threads x 8 columns; every iteration takes EH_R rows.  The per-row logits go warp-shuffle ->
// shared (double-buffered by iteration parity: ONE __syncthreads per iteration) and every warp then
// finishes the EH_R rows redundantly in its first 2*EH_R lanes (lane = 2*row + output), so no thread
// waits on a serial tail; warp 0 alone writes the row outputs.
static __global__ void __launch_bounds__(256, 2) k_edge_head_fwd(const EdgeHeadP a) {
  pdl_wait();
  __shared__ float red[2][8][2 * EH_R];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int c0 = t * 8;
  float w0[8], w1[8], acc0[8], acc1[8];
  {
    const float4 x0 = ld4f(a.W2 + c0), x1 = ld4f(a.W2 + c0 + 4), y0 = ld4f(a.W2 + 4 * H + c0), y1 = ld4f(a.W2 + 4 * H + c0 + 4);
    w0[0] = x0.x; w0[1] = x0.y; w0[2] = x0.z; w0[3] = x0.w; w0[4] = x1.x; w0[5] = x1.y; w0[6] = x1.z; w0[7] = x1.w;
    w1[0] = y0.x; w1[1] = y0.y; w1[2] = y0.z; w1[3] = y0.w; w1[4] = y1.x; w1[5] = y1.y; w1[6] = y1.z; w1[7] = y1.w;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
  float accb[8];                                       // the head's pre-activation gradient summed over rows (h_to_edge.0.bias)
#pragma unroll
  for (int k = 0; k < 8; ++k) accb[k] = 0.f;
  float dbsum = 0.f;                                   // lanes 0..2*EH_R-1 of warp 0: sum of their dl
  const float bias = a.b2[lane & 1];
  int par = 0;
  for (int r0 = blockIdx.x * EH_R; r0 < a.B; r0 += gridDim.x * EH_R, par ^= 1) {
    // target flag of (row lane>>1, output lane&1): issued before the big loads, used after the reduction
    const int rb = r0 + (lane >> 1);
    float tgt = 0.f;
    const bool greedy = a.adj_out != nullptr;
    if (!greedy && lane < 2 * EH_R && rb < a.B) {
      const uint64_t A = a.adj[rb];
      tgt = (lane & 1) ? (float)abit(A, a.vi, a.vj) : (float)abit(A, a.vj, a.vi);
    }
    // all 4*EH_R 128-bit loads of the iteration are issued before anything consumes them (rows past the end
    // re-read the last row: their dl is forced to 0 and their mask is not stored)
    float4 lu[EH_R][2], lq[EH_R][2];
#pragma unroll
    for (int r = 0; r < EH_R; ++r) {
      const int bb = min(r0 + r, a.B - 1);
      const float* up = a.U + (int64_t)bb * 4 * H + c0; const float* qp = a.Q + (int64_t)bb * 4 * H + c0;
      lu[r][0] = ld4f(up); lu[r][1] = ld4f(up + 4); lq[r][0] = ld4f(qp); lq[r][1] = ld4f(qp + 4);
    }
    float e[EH_R][8], part[2 * EH_R];
    unsigned mk[EH_R];
#pragma unroll
    for (int r = 0; r < EH_R; ++r) {
      const float4 u0 = lu[r][0], u1 = lu[r][1], q0 = lq[r][0], q1 = lq[r][1];
      e[r][0] = fmaxf(u0.x + q0.x, 0.f); e[r][1] = fmaxf(u0.y + q0.y, 0.f); e[r][2] = fmaxf(u0.z + q0.z, 0.f);
      e[r][3] = fmaxf(u0.w + q0.w, 0.f); e[r][4] = fmaxf(u1.x + q1.x, 0.f); e[r][5] = fmaxf(u1.y + q1.y, 0.f);
      e[r][6] = fmaxf(u1.z + q1.z, 0.f); e[r][7] = fmaxf(u1.w + q1.w, 0.f);
      float p0 = 0.f, p1 = 0.f; unsigned m = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { p0 = fmaf(e[r][k], w0[k], p0); p1 = fmaf(e[r][k], w1[k], p1); m |= (e[r][k] > 0.f ? 1u : 0u) << k; }
      mk[r] = m;
      part[2 * r] = p0; part[2 * r + 1] = p1;
    }
#pragma unroll
    for (int r = 0; r < EH_R; ++r)
      if (a.mask && r0 + r < a.B) a.mask[(int64_t)(r0 + r) * 256 + t] = (uint8_t)mk[r];
#pragma unroll
    for (int i = 0; i < 2 * EH_R; ++i) {
      float v = part[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[par][wid][i] = v;
    }
    __syncthreads();
    float dl = 0.f;
    if (lane < 2 * EH_R) {
      float l = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) l += red[par][w][lane];
      l += bias;
      const bool ok = rb < a.B;
      if (greedy) {
        if (wid == 0) {
          const float l_other = __shfl_down_sync((1u << (2 * EH_R)) - 1u, l, 1);     // the row's second logit (odd lane)
          if (ok && !(lane & 1)) edge_head_row_decide(a, rb, l, l_other);
        }
      } else {
      dl = ok ? (sigmoidf_(l) - tgt) * a.inv_batch : 0.f;
      if (wid == 0) {
        float bce = ok ? bce_logits(l, tgt) : 0.f;
        bce += __shfl_down_sync((1u << (2 * EH_R)) - 1u, bce, 1);
        if (ok) {
          a.l2[(int64_t)rb * LD_E + (lane & 1)] = l; a.dl2[(int64_t)rb * LD_E + (lane & 1)] = dl;
          if (!(lane & 1)) a.rowloss[(int64_t)2 * a.B + rb] += bce * a.inv_batch;
        }
        dbsum += dl;
      }
      }
    }
    if (a.dW2) {
#pragma unroll
      for (int r = 0; r < EH_R; ++r) {
        const float d0 = __shfl_sync(0xffffffffu, dl, 2 * r), d1 = __shfl_sync(0xffffffffu, dl, 2 * r + 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc0[k] = fmaf(d0, e[r][k], acc0[k]); acc1[k] = fmaf(d1, e[r][k], acc1[k]); }
        if (a.db0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) accb[k] += e[r][k] > 0.f ? d0 * w0[k] + d1 * w1[k] : 0.f;
        }
      }
    }
  }
  if (a.dW2) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(a.dW2 + c0 + k, acc0[k]); atomicAdd(a.dW2 + 4 * H + c0 + k, acc1[k]); }
    if (a.db0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(a.db0 + c0 + k, accb[k]);
    }
    if (wid == 0) {                                    // db2[c] = sum over rows: fold lanes of equal parity
      float v = dbsum;
      v += __shfl_down_sync(0xffffffffu, v, 4); v += __shfl_down_sync(0xffffffffu, v, 2);
      if (lane < 2) atomicAdd(a.db2 + lane, v);
    }
  }
}
static void edge_head_fwd(dx_stream_t st, const EdgeHeadP& a) {
  int blocks = (a.B + EH_R - 1) / EH_R;
  if (blocks > 148 * 2) blocks = 148 * 2;
  launch_k(k_edge_head_fwd, dim3(blocks), dim3(256), 0, st, 1, a);
  ++g_launches;
}
#else
static void edge_head_fwd(dx_stream_t, const EdgeHeadP& a) {
  for (int b = 0; b < a.B; ++b) {
    float l0 = 0.f, l1 = 0.f;
    for (int j = 0; j < 4 * H; ++j) {
      const float e = fmaxf(a.U[(int64_t)b * 4 * H + j] + a.Q[(int64_t)b * 4 * H + j], 0.f);
      l0 += e * a.W2[j]; l1 += e * a.W2[4 * H + j];
      if (a.mask) {
        if ((j & 7) == 0) a.mask[(int64_t)b * 256 + j / 8] = 0;
        if (e > 0.f) a.mask[(int64_t)b * 256 + j / 8] |= (uint8_t)(1u << (j & 7));
      }
    }
    if (a.adj_out) { edge_head_row_decide(a, b, l0 + a.b2[0], l1 + a.b2[1]); continue; }
    float dl[2];
    edge_head_row_finish(a, b, l0, l1, dl);
    if (a.dW2) {
      for (int j = 0; j < 4 * H; ++j) {
        const float e = fmaxf(a.U[(int64_t)b * 4 * H + j] + a.Q[(int64_t)b * 4 * H + j], 0.f);
        a.dW2[j] += dl[0] * e; a.dW2[4 * H + j] += dl[1] * e;
        if (a.db0 && e > 0.f) a.db0[j] += dl[0] * a.W2[j] + dl[1] * a.W2[4 * H + j];
      }
      a.db2[0] += dl[0]; a.db2[1] += dl[1];
    }
  }
  ++g_launches;
}
#endif
// Backward of the heads w.r.t. their pre-activations.  A head gradient is rank-2 under its relu mask,
//   g_t[b,:] = mask_t[b,:] * (dl0_t[b] W2[0,:] + dl1_t[b] W2[1,:]),
// so sums of them (dQ of a finished node over its later readers; dU of a node state over the heads that
// read that state) are re-formed from the 256 B/row masks where they are consumed instead of being
// accumulated through HBM (8 KB/row read-modify-write per step and buffer).
//   out[m,:] = sum_k g_{step k}[b,:]   over the listed steps, b = rows ? rows[m] : m ;
//   until_active: stop after the first listed step at which graph b adds an edge to node vi
//   (that step replaces the state, so earlier-listed heads are the only readers of this version).
struct HeadSumP {
  int M; const int* rows; const uint64_t* adj; int vi; int n; int until_active;
  const uint8_t* mask[6]; const float* dl[6]; int vj[6];
  const float* W2; float* out;
};
DX_HD DX_INLINE void head_sum_item(const HeadSumP& a, int m, int tc, const float* w0, const float* w1) {
  const int64_t b = a.rows ? a.rows[m] : m;
  const uint64_t A = a.adj[b];
  float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  // the (at most 6) mask bytes and dl pairs are loaded up front: independent loads, no serial chain
  unsigned mk[6]; float d0[6], d1[6];
#ifndef DX_EMU
#pragma unroll
#endif
  for (int k = 0; k < 6; ++k) {
    const int kk = k < a.n ? k : 0;
    mk[k] = a.mask[kk][b * 256 + tc];
    const float2 dd = *reinterpret_cast<const float2*>(a.dl[kk] + b * LD_E);
    d0[k] = dd.x; d1[k] = dd.y;
  }
  bool live = true;
#ifndef DX_EMU
#pragma unroll
#endif
  for (int k = 0; k < 6; ++k) {
    if (k < a.n && live) {
      for (int e = 0; e < 8; ++e) g[e] += ((mk[k] >> e) & 1u) ? d0[k] * w0[e] + d1[k] * w1[e] : 0.f;
      if (a.until_active && (abit(A, a.vj[k], a.vi) | abit(A, a.vi, a.vj[k]))) live = false;
    }
  }
  float* o = a.out + (int64_t)m * 4 * H + tc * 8;
  st4f(o, make_float4(g[0], g[1], g[2], g[3])); st4f(o + 4, make_float4(g[4], g[5], g[6], g[7]));
}
#ifndef DX_EMU
// thread = one group of 8 columns (its two W2 slices stay in registers), blocks stride over the rows
static __global__ void __launch_bounds__(256, 4) k_head_sum(const HeadSumP a) {
  pdl_wait();
  const int tc = threadIdx.x, c0 = tc * 8;
  float w0[8], w1[8];
  {
    const float4 x0 = ld4f(a.W2 + c0), x1 = ld4f(a.W2 + c0 + 4), y0 = ld4f(a.W2 + 4 * H + c0), y1 = ld4f(a.W2 + 4 * H + c0 + 4);
    w0[0] = x0.x; w0[1] = x0.y; w0[2] = x0.z; w0[3] = x0.w; w0[4] = x1.x; w0[5] = x1.y; w0[6] = x1.z; w0[7] = x1.w;
    w1[0] = y0.x; w1[1] = y0.y; w1[2] = y0.z; w1[3] = y0.w; w1[4] = y1.x; w1[5] = y1.y; w1[6] = y1.z; w1[7] = y1.w;
  }
  for (int m = blockIdx.x; m < a.M; m += gridDim.x) head_sum_item(a, m, tc, w0, w1);
}
static void head_sum(dx_stream_t st, const HeadSumP& a) {
  if (a.M <= 0) return;
  launch_k(k_head_sum, dim3(a.M < 148 * 8 ? a.M : 148 * 8), dim3(256), 0, st, 1, a);
  ++g_launches;
}
#else
static void head_sum(dx_stream_t, const HeadSumP& a) {
  for (int m = 0; m < a.M; ++m)
    for (int tc = 0; tc < 256; ++tc) head_sum_item(a, m, tc, a.W2 + tc * 8, a.W2 + 4 * H + tc * 8);
  ++g_launches;
}
#endif

// y[m] = a[m, :] . w + b for a one-output head (h_to_edge_self.2: N = 1, K = 1024).  A GEMM tile wastes 63/64 of its
// columns and, at small batches, runs on two CTAs; here one warp takes a row (128-bit loads, shuffle reduction).
#ifndef DX_EMU
static __global__ void __launch_bounds__(256) k_rowdot(int M, int K, const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ out, int64_t ldo) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), nw = (int)((gridDim.x * (int64_t)blockDim.x) >> 5);
  for (int m = warp; m < M; m += nw) {
    const float* a = A + (int64_t)m * lda;
    float s = 0.f;
#pragma unroll 8
    for (int k = lane * 4; k < K; k += 128) {
      const float4 x = ld4f(a + k), y = ld4f(w + k);
      s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[(int64_t)m * ldo] = s + (b ? b[0] : 0.f);
  }
}
static void rowdot(dx_stream_t st, int M, int K, const float* A, int64_t lda, const float* w, const float* b, float* out, int64_t ldo) {
  int blocks = (M + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(k_rowdot, dim3(blocks), dim3(256), 0, st, 1, M, K, A, lda, w, b, out, ldo);
  ++g_launches;
}
#else
static void rowdot(dx_stream_t, int M, int K, const float* A, int64_t lda, const float* w, const float* b, float* out, int64_t ldo) {
  for (int m = 0; m < M; ++m) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(A[(int64_t)m * lda + k], w[k], s);
    out[(int64_t)m * ldo] = s + (b ? b[0] : 0.f);
  }
  ++g_launches;
}
#endif

// =============================================================================================
// forward
// =============================================================================================

static void mlp3_fwd(dx_stream_t st, const Weights& W, int B, int w0, const float* hin, int nout, float* A1, float* A2,
                     float* L) {
  linear_fwd(st, B, 2 * H, H, hin, H, W[w0], H, W[w0 + 1], A1, 2 * H, ACT_RELU);
  linear_fwd(st, B, 2 * H, 2 * H, A1, 2 * H, W[w0 + 2], 2 * H, W[w0 + 3], A2, 2 * H, ACT_RELU);
  linear_fwd(st, B, nout, 2 * H, A2, 2 * H, W[w0 + 4], 2 * H, W[w0 + 5], L, LD_L);
}

// bt != NULL (teacher forcing with the step schedule): a later node vi reads the "out" half of node v's projections
// through its forward edge vi -> v and the "in" half only through a feedback back-edge v -> vi, so the "in" half is
// computed on the rows of schedule list NSTEP+6+v alone (gather -> product -> scatter).  Greedy decoding does not know
// its edges yet and computes both halves for every graph.
static void node_projections(dx_stream_t st, const Weights& W, int B, int v, const DecWs& w, const Batch* bt = nullptr) {
  float* h = w.Hd + (size_t)v * B * H;
  float* Pg = w.Pg + (size_t)v * B * 2 * H; float* Pm = w.Pm + (size_t)v * B * 2 * H;
  proj_fwd(st, B, h, W[P_G_W], Pg, HALF_OUT);
  proj_fwd(st, B, h, W[P_M_W], Pm, HALF_OUT);
  if (bt && bt->step_ptr) {
    const int tl = NSTEP + 6 + v, n = bt->step_ptr[tl + 1] - bt->step_ptr[tl];
    if (n > 0) {
      const int* rows = bt->step_rows + bt->step_ptr[tl];
      float* th = w.UC; float* tg = w.UC + (size_t)B * H; float* tm = w.UC + (size_t)2 * B * H;   // (UC is idle between steps)
      gather_rows(st, n, H, rows, h, th, 0);
      linear_fwd(st, n, H, H, th, H, W[P_G_W], 2 * H, nullptr, tg, H);
      linear_fwd(st, n, H, H, th, H, W[P_M_W], 2 * H, nullptr, tm, H);
      scatter_rows(st, n, H, rows, tg, Pg, 0, 2 * H);
      scatter_rows(st, n, H, rows, tm, Pm, 0, 2 * H);
    }
  } else {
    proj_fwd(st, B, h, W[P_G_W], Pg, HALF_IN);
    proj_fwd(st, B, h, W[P_M_W], Pm, HALF_IN);
  }
  // Hj half of h_to_edge.0 (columns 512..1023 of the (2048,1024) weight) + its bias
  linear_fwd(st, B, 4 * H, H, h, H, W[P_E_W0] + H, 2 * H, W[P_E_B0], w.Q + (size_t)v * B * 4 * H, 4 * H);
}

void decode_fwd_impl(dx_stream_t st, const Weights& W, int B, const float* z, const DecWs& w, const DecIO& io) {
  const bool train = io.train;
  const uint64_t* adj = train ? io.bt->adj : io.adj_out;
  const float* Xsrc = train ? io.bt->Xn : w.Xd;   // node-major (7,B,32)
  upload_tables();
  const float* Wc = W[P_CD_WIH]; const float* Wl = W[P_LD_WIH]; const float* Wr = W[P_RD_WIH];
  int Kx = SX, Kr = SX0, ldx = SX, ldr = SX0;
  if (train || get_precision() != PREC_FP32) {   // 32-column padded input weights: TMA-addressable (exact: the padded columns are zero)
    pad_wih(st, W[P_CD_WIH], SX, w.WihP[0]); pad_wih(st, W[P_LD_WIH], SX, w.WihP[1]); pad_wih(st, W[P_RD_WIH], SX0, w.WihP[2]);
    Wc = w.WihP[0]; Wl = w.WihP[1]; Wr = w.WihP[2]; Kx = Kr = ldx = ldr = XP;
  }

  linear_fwd(st, B, H, Z, z, Z, W[P_ZH_W], Z, W[P_ZH_B], w.Hinit, H, ACT_TANH);
  mlp3_fwd(st, W, B, P_X0_W0, w.Hinit, SX0 + 32, w.A1[0], w.A2[0], w.L[0]);
  if (train) loss_x0(st, B, w.L[0], Xsrc, io.bt->cls, io.lw, w.rowloss, w.dL[0]);
  else reg_x0(st, B, w.L[0], w.Xd, w.Pn, io.margins);
  // root: h_0 = GRU_root(x0[:23], H_init)
  linear_fwd(st, B, G3, Kr, Xsrc, XP, Wr, ldr, nullptr, w.gxc[0], G3);
  linear_fwd(st, B, G3, H, w.Hinit, H, W[P_RD_WHH], H, nullptr, w.gh, G3);
  {
    RowMap rm{B, B, nullptr, 0};
    CellFwd c{rm, w.gxc[0], w.gh, W[P_RD_BIH], W[P_RD_BHH], w.Hinit, 0, w.Hd, 0, train ? w.g_root : nullptr, 0, S_ONE, adj};
    cell_fwd(st, c);
  }
  node_projections(st, W, B, 0, w, train ? io.bt : nullptr);

  const bool p1b = p1_batched(B, train, train ? io.bt->step_ptr : nullptr);
  if (p1b) {
    // first propagates of nodes 1..6 in one pass over 6B rows (see p1_batched): row m of every buffer below is graph
    // m % B of node 1 + m / B
    const int B6 = 6 * B;
    const float* X1 = Xsrc + (size_t)B * XP;
    RowMap rm6{B6, B, nullptr, B};
    linear_fwd(st, B6, G3, Kx, X1, XP, Wc, ldx, nullptr, w.gxc[1], G3);
    linear_fwd(st, B6, G3, Kx, X1, XP, Wl, ldx, nullptr, w.gxl[1], G3);
    CellFwd c0{rm6, w.gxc[1], nullptr, W[P_CD_BIH], W[P_CD_BHH], nullptr, 0, w.Hc0[1], 0, w.g_c0[1], 0, S_ONE, adj};
    cell_fwd(st, c0);
    linear_fwd(st, B6, G3, H, w.Hc0[1], H, W[P_LD_WHH], H, nullptr, w.ghl0, G3);
    CellFwd p1{rm6, w.gxl[1], w.ghl0, W[P_LD_BIH], W[P_LD_BHH], w.Hc0[1], 0, w.Hi_p1[1], 0, w.g_p1[1], 0, S_ZERO, adj};
    p1.hout2 = w.Hi_p2[1]; p1.hout3 = w.Hd + (size_t)B * H; p1.hout23_local = 1;   // P2 state and current state start as P1's
    cell_fwd(st, p1);
    linear_fwd(st, B6, 2 * H, H, w.Hi_p1[1], H, W[P_ES_W0], H, W[P_ES_B0], w.ES1[1], 2 * H, ACT_RELU);
    rowdot(st, B6, 2 * H, w.ES1[1], 2 * H, W[P_ES_W2], W[P_ES_B2], w.ls[1], LD_E);
    // second propagates (x_loop = s*x): the self-loop rows of each node (schedule list NSTEP+vi-1), gates stored compactly
    for (int vi = 1; vi < NN; ++vi) {
      const int ts = NSTEP + vi - 1, ns = io.bt->step_ptr[ts + 1] - io.bt->step_ptr[ts];
      if (ns <= 0) continue;
      RowMap rs{ns, B, io.bt->step_rows + io.bt->step_ptr[ts], vi * B};
      CellFwd p2{rs, w.gxl[vi], w.ghl0 + (size_t)(vi - 1) * B * G3, W[P_LD_BIH], W[P_LD_BHH], w.Hc0[vi], 0, w.UC, 0, w.g_p2[vi], 0,
                 S_SELF, adj};
      p2.gx_by_graph = 1; p2.gh_by_graph = 1; p2.hprev_by_graph = 1; p2.hout2 = w.Hi_p2[vi]; p2.hout3 = w.Hd + (size_t)vi * B * H;
      cell_fwd(st, p2);
    }
    // U = Hi_p2 W_e0[:, :512]^T of every node: the running edge-head product each node's steps start from
    linear_fwd(st, B6, 4 * H, H, w.Hi_p2[1], H, W[P_E_W0], 2 * H, nullptr, w.U6, 4 * H);
  }
  int t = 0;
  for (int vi = 1; vi < NN; ++vi) {
    const float* hprev_node = w.Hd + (size_t)(vi - 1) * B * H;
    float* Xi = const_cast<float*>(Xsrc) + (size_t)vi * B * XP;
    RowMap rm{B, B, nullptr, vi * B};
    if (!train) {
      mlp3_fwd(st, W, B, P_X_W0, hprev_node, SX, w.A1[vi], w.A2[vi], w.L[vi]);
      reg_xi(st, B, w.L[vi], w.Xd + (size_t)vi * B * XP, w.Pn + (size_t)vi * B * XP, io.margins);
    } else if (!heads_batched(B, train)) {                         // (batched heads: one pass after the node loop)
      mlp3_fwd(st, W, B, P_X_W0, hprev_node, SX, w.A1[vi], w.A2[vi], w.L[vi]);
      loss_xi(st, B, vi, w.L[vi], Xi, io.bt->cls, io.lw, w.rowloss, w.dL[vi]);
    }
    const float* const ghl0 = p1b ? w.ghl0 + (size_t)(vi - 1) * B * G3 : w.ghl0;
    const bool compact = train && io.bt->step_ptr != nullptr;
    float* const Hcur = w.Hd + (size_t)vi * B * H;   // compacted steps / greedy: the CURRENT state of node vi for every graph
    if (!p1b) {
    linear_fwd(st, B, G3, Kx, Xi, XP, Wc, ldx, nullptr, w.gxc[vi], G3);
    linear_fwd(st, B, G3, Kx, Xi, XP, Wl, ldx, nullptr, w.gxl[vi], G3);
    // P1 (model.py:234/320): no edges yet -> H_in = 0, x_loop = 0
    CellFwd c0{rm, w.gxc[vi], nullptr, W[P_CD_BIH], W[P_CD_BHH], nullptr, 0, w.Hc0[vi], 0, train ? w.g_c0[vi] : nullptr, 0,
               S_ONE, adj};
    cell_fwd(st, c0);
    linear_fwd(st, B, G3, H, w.Hc0[vi], H, W[P_LD_WHH], H, nullptr, w.ghl0, G3);
    CellFwd p1{rm, w.gxl[vi], w.ghl0, W[P_LD_BIH], W[P_LD_BHH], w.Hc0[vi], 0, w.Hi_p1[vi], 0, train ? w.g_p1[vi] : nullptr,
               0, S_ZERO, adj};
    // compacted steps: P2 equals P1 except on the self-loop rows, so P1 also fills the P2 state and the current state
    // (the self-loop rows are overwritten below) instead of two device copies
    if (compact) { p1.hout2 = w.Hi_p2[vi]; p1.hout3 = Hcur; }
    cell_fwd(st, p1);
    // self-loop head (model.py:236/331)
    linear_fwd(st, B, 2 * H, H, w.Hi_p1[vi], H, W[P_ES_W0], H, W[P_ES_B0], w.ES1[vi], 2 * H, ACT_RELU);
    rowdot(st, B, 2 * H, w.ES1[vi], 2 * H, W[P_ES_W2], W[P_ES_B2], w.ls[vi], LD_E);
    }
    // (batched first propagates address ls / dls as ONE [6B, LD_E] matrix from the node-1 buffer: 16-byte rows, so the
    // arena's 256-byte alignment may pad between the per-node pointers)
    if (train) loss_edge(st, B, vi, vi, p1b ? w.ls[1] + (size_t)(vi - 1) * B * LD_E : w.ls[vi], 1, adj, io.lw, w.rowloss,
                         p1b ? w.dls[1] + (size_t)(vi - 1) * B * LD_E : w.dls[vi]);
    else decide_edges(st, B, vi, vi, w.ls[vi], 1, io.adj_out, io.margins);
    // P2 (model.py:240/337): same H_in = 0, x_loop = s*x
    if (p1b) {
      // (done for all nodes ahead of the loop)
    } else if (compact) {
      // x_loop = s*x: P2 repeats P1 exactly on graphs without a self-loop on vi, so it is computed on the
      // self-loop rows only (schedule list NSTEP+vi-1); their gates are stored compactly.
      const int ts = NSTEP + vi - 1, ns = io.bt->step_ptr[ts + 1] - io.bt->step_ptr[ts];
      if (ns > 0) {
        RowMap rs{ns, B, io.bt->step_rows + io.bt->step_ptr[ts], vi * B};
        CellFwd p2{rs, w.gxl[vi], ghl0, W[P_LD_BIH], W[P_LD_BHH], w.Hc0[vi], 0, w.UC, 0, w.g_p2[vi], 0, S_SELF, adj};
        p2.gx_by_graph = 1; p2.gh_by_graph = 1; p2.hprev_by_graph = 1; p2.hout2 = w.Hi_p2[vi]; p2.hout3 = Hcur;
        cell_fwd(st, p2);
      }
    } else {
      CellFwd p2{rm, w.gxl[vi], w.ghl0, W[P_LD_BIH], W[P_LD_BHH], w.Hc0[vi], 0, w.Hi_p2[vi], 0, train ? w.g_p2[vi] : nullptr,
                 0, S_SELF, adj};
      if (!train) p2.hout2 = Hcur;                         // greedy: the current state starts as the P2 state
      cell_fwd(st, p2);
    }
    if (train && !compact) zero_async(st, w.Hrun, sizeof(float) * (size_t)B * H);   // (compacted / greedy steps: first-touch, msg_fwd accum = 2)
    if (compact) {
      // Compacted teacher forcing (DESIGN.md "identity steps"): a re-propagate changes node vi only for
      // graphs where the step adds an edge.  Hd[vi] holds the CURRENT state of node vi for every
      // graph and U = Hd[vi] W_e0[:, :512]^T is kept consistent with it, so the edge head of a step is
      // element-wise for all graphs and the GRU / projection products run on the active rows only.
      float* const Ucur = p1b ? w.U6 + (size_t)(vi - 1) * B * 4 * H : w.U;
      if (!p1b) linear_fwd(st, B, 4 * H, H, w.Hi_p2[vi], H, W[P_E_W0], 2 * H, nullptr, Ucur, 4 * H);
      for (int vj = vi - 1; vj >= 0; --vj, ++t) {
        // fused edge head: the E1 buffer of the step only stores the relu bit-mask (256 B/row)
        EdgeHeadP eh{B, vi, vj, Ucur, w.Q + (size_t)vj * B * 4 * H, W[P_E_W2], W[P_E_B2], adj, io.lw.inv_batch, w.l2[t], w.dl2[t],
                     w.rowloss, reinterpret_cast<uint8_t*>(w.E1[t]), io.dW2, io.db2};
        eh.db0 = io.db0;
        edge_head_fwd(st, eh);
        const int n = io.bt->step_ptr[t + 1] - io.bt->step_ptr[t];
        if (n <= 0) continue;
        const int* rows = io.bt->step_rows + io.bt->step_ptr[t];
        RowMap rc{n, B, rows, vi * B};
        MsgFwd mf{rc, w.Pg, w.Pm, W[P_G_B], adj, w.Hrun, 0, vj, vj, 2};
        mf.hin_by_graph = 1; mf.hin_copy = w.Hin[t];
        msg_fwd(st, mf);
        linear_fwd(st, n, G3, H, w.Hin[t], H, W[P_CD_WHH], H, nullptr, w.gh, G3);
        CellFwd cc{rc, w.gxc[vi], w.gh, W[P_CD_BIH], W[P_CD_BHH], w.Hin[t], 0, w.Hc[t], 0, w.g_c[t], 0, S_ONE, adj};
        cc.gx_by_graph = 1;
        cell_fwd(st, cc);
        linear_fwd(st, n, G3, H, w.Hc[t], H, W[P_LD_WHH], H, nullptr, w.gh, G3);
        CellFwd cl{rc, w.gxl[vi], w.gh, W[P_LD_BIH], W[P_LD_BHH], w.Hc[t], 0, w.Hi[t], 0, w.g_l[t], 0, S_SELF, adj};
        cl.gx_by_graph = 1; cl.hout2 = Hcur;
        cell_fwd(st, cl);
        if (vj > 0) {                                          // (no head reads U after the node's last step)
          linear_fwd(st, n, 4 * H, H, w.Hi[t], H, W[P_E_W0], 2 * H, nullptr, w.UC, 4 * H);
          scatter_rows(st, n, 4 * H, rows, w.UC, Ucur, 0);
        }
      }
      if (vi < NN - 1) node_projections(st, W, B, vi, w, io.bt);
      continue;
    }
    if (!train) {
      // Greedy generation with the same identity-step compaction: after each pair of edge decisions only the
      // graphs that gained an edge re-propagate (the others would recompute the state they already have).  The
      // active rows are compacted on the device and their count read back (one small sync per step) to size the
      // products; U = Hd[vi] W_e0[:, :512]^T is kept current, so the edge head is element-wise for all graphs.
      linear_fwd(st, B, 4 * H, H, w.Hi_p2[vi], H, W[P_E_W0], 2 * H, nullptr, w.U, 4 * H);
      for (int vj = vi - 1; vj >= 0; --vj, ++t) {
        EdgeHeadP eh{B, vi, vj, w.U, w.Q + (size_t)vj * B * 4 * H, W[P_E_W2], W[P_E_B2], nullptr, 0.f, w.l2[t], nullptr,
                     nullptr, nullptr, nullptr, nullptr};
        eh.adj_out = io.adj_out; eh.margins = io.margins; eh.active = w.act_flag;
        edge_head_fwd(st, eh);
        const int n = compact_flags(st, B, w.act_flag, w.act_rows, w.act_cnt);
        if (n <= 0) continue;
        RowMap rc{n, B, w.act_rows, vi * B};
        float* HinC = w.Hin[t]; float* HcC = w.Hc[t]; float* HiC = w.Hi[t];       // (one shared scratch each when not training)
        MsgFwd mf{rc, w.Pg, w.Pm, W[P_G_B], adj, w.Hrun, 0, vj, vj, 2};
        mf.hin_by_graph = 1; mf.hin_copy = HinC;
        msg_fwd(st, mf);
        linear_fwd(st, n, G3, H, HinC, H, W[P_CD_WHH], H, nullptr, w.gh, G3);
        CellFwd cc{rc, w.gxc[vi], w.gh, W[P_CD_BIH], W[P_CD_BHH], HinC, 0, HcC, 0, nullptr, 0, S_ONE, adj};
        cc.gx_by_graph = 1;
        cell_fwd(st, cc);
        linear_fwd(st, n, G3, H, HcC, H, W[P_LD_WHH], H, nullptr, w.gh, G3);
        CellFwd cl{rc, w.gxl[vi], w.gh, W[P_LD_BIH], W[P_LD_BHH], HcC, 0, HiC, 0, nullptr, 0, S_SELF, adj};
        cl.gx_by_graph = 1; cl.hout2 = Hcur;
        cell_fwd(st, cl);
        if (vj > 0) {                                          // (no head reads U after the node's last step)
          linear_fwd(st, n, 4 * H, H, HiC, H, W[P_E_W0], 2 * H, nullptr, w.UC, 4 * H);
          scatter_rows(st, n, 4 * H, w.act_rows, w.UC, w.U, 0);
        }
      }
      if (vi < NN - 1) node_projections(st, W, B, vi, w);
      continue;
    }
    const float* Hi_prev = w.Hi_p2[vi];
    for (int vj = vi - 1; vj >= 0; --vj, ++t) {
      // edge head on cat[Hi, Hj] (model.py:245/350): first layer split into Hi half + cached Hj half
      linear_fwd(st, B, 4 * H, H, Hi_prev, H, W[P_E_W0], 2 * H, nullptr, w.E1[t], 4 * H, ACT_RELU, nullptr, nullptr,
                 w.Q + (size_t)vj * B * 4 * H, 4 * H);
      linear_fwd(st, B, 2, 4 * H, w.E1[t], 4 * H, W[P_E_W2], 4 * H, W[P_E_B2], w.l2[t], LD_E);
      if (train) loss_edge(st, B, vi, vj, w.l2[t], 2, adj, io.lw, w.rowloss, w.dl2[t]);
      else decide_edges(st, B, vi, vj, w.l2[t], 2, io.adj_out, io.margins);
      // add the message of vj to the running aggregate, then re-propagate (model.py:251/358)
      MsgFwd mf{rm, w.Pg, w.Pm, W[P_G_B], adj, w.Hrun, 0, vj, vj, 1};
      msg_fwd(st, mf);
      if (train) copy_async(st, w.Hin[t], w.Hrun, sizeof(float) * (size_t)B * H);
      const float* Hin_t = train ? w.Hin[t] : w.Hrun;
      linear_fwd(st, B, G3, H, Hin_t, H, W[P_CD_WHH], H, nullptr, w.gh, G3);
      CellFwd cc{rm, w.gxc[vi], w.gh, W[P_CD_BIH], W[P_CD_BHH], Hin_t, 0, w.Hc[t], 0, train ? w.g_c[t] : nullptr, 0, S_ONE,
                 adj};
      cell_fwd(st, cc);
      linear_fwd(st, B, G3, H, w.Hc[t], H, W[P_LD_WHH], H, nullptr, w.gh, G3);
      float* Hi_out = (vj == 0) ? w.Hd + (size_t)vi * B * H : w.Hi[t];
      CellFwd cl{rm, w.gxl[vi], w.gh, W[P_LD_BIH], W[P_LD_BHH], w.Hc[t], 0, Hi_out, 0, train ? w.g_l[t] : nullptr, 0, S_SELF,
                 adj};
      cell_fwd(st, cl);
      Hi_prev = Hi_out;
    }
    if (vi < NN - 1) node_projections(st, W, B, vi, w);
  }
  if (heads_batched(B, train)) {
    // parameter heads of nodes 1..6 (model.py:318/323-328) on the finished states h_0..h_5: Hd, A1, A2, L, dL are
    // node-major and carved back to back, so rows [0,6B) of Hd map to rows [0,6B) of the node-1 buffers
    mlp3_fwd(st, W, 6 * B, P_X_W0, w.Hd, SX, w.A1[1], w.A2[1], w.L[1]);
    loss_xi_all(st, B, w.L[1], Xsrc, io.bt->cls, io.lw, w.rowloss, w.dL[1]);
  }
}

// =============================================================================================
// backward (teacher-forced loss only)
// =============================================================================================
static void mlp3_bwd(dx_stream_t st, const Weights& W, const Weights& G, int B, int w0, const float* hin, int nout,
                     const float* A1, const float* A2, const float* dL, const DecWs& w, float* dhin, int dhin_accum = ACC_ADD) {
  linear_wgrad(st, B, nout, 2 * H, dL, LD_L, A2, 2 * H, G[w0 + 4], 2 * H);
  colsum_accum(st, B, nout, dL, LD_L, G[w0 + 5]);
  linear_dgrad(st, B, nout, 2 * H, dL, LD_L, W[w0 + 4], 2 * H, w.dA2, 2 * H, ACC_STORE, nullptr, nullptr, A2, 2 * H);   // relu backward fused
  linear_wgrad(st, B, 2 * H, 2 * H, w.dA2, 2 * H, A1, 2 * H, G[w0 + 2], 2 * H);
  colsum_accum(st, B, 2 * H, w.dA2, 2 * H, G[w0 + 3]);
  linear_dgrad(st, B, 2 * H, 2 * H, w.dA2, 2 * H, W[w0 + 2], 2 * H, w.dA1, 2 * H, ACC_STORE, nullptr, nullptr, A1, 2 * H);
  linear_wgrad(st, B, 2 * H, H, w.dA1, 2 * H, hin, H, G[w0], H);
  colsum_accum(st, B, 2 * H, w.dA1, 2 * H, G[w0 + 1]);
  linear_dgrad(st, B, 2 * H, H, w.dA1, 2 * H, W[w0], H, dhin, H, dhin_accum);
}

// looper cell backward for one propagate: dHi -> (dHc += ..., weight grads)
// x_stash != NULL (a compacted step): the gate gradients are written IN PLACE over the step's saved gates (each thread
// reads its four gate values before it stores the four gradients at the same addresses; nothing reads the gates
// again) and the masked x rows go to x_stash — both weight gradients of the cell are then formed ONCE for all steps at
// the end of decode_bwd_impl from the concatenated buffers instead of two products per step.
static void looper_bwd(dx_stream_t st, const Weights& W, const Weights& G, int B, int vi, const RowMap& rm,   // rm.M rows
                       const float* dHi, const float* gates, const float* Hc, int smode, const uint64_t* adj,
                       const float* Xi, const DecWs& w, float* dHc, bool dHc_accum, float* x_stash = nullptr,
                       bool defer = false, bool x_stashed = false) {   // x_stashed: stash_step_x already filled x_stash                        // defer without x_stash: the S_ZERO propagate (no x term)
  // dHc (+)= dHi*z + dgh W_hh            (rm.M rows: all B graphs, or the active rows of a compacted step)
  const int M = rm.M;
  float* direct = dHc_accum ? w.dHin : dHc;  // dHin is free scratch at this point
  defer = defer || x_stash != nullptr;
  float* d4 = defer ? const_cast<float*>(gates) : w.dgx;
  float* dgh = d4 + H;
  CellBwd cb{rm, dHi, 0, gates, 0, Hc, 0, d4, nullptr, dgh, direct, smode, adj};
  cell_bwd(st, cb, G[P_LD_BIH], G[P_LD_BHH]);
  if (dHc_accum) add_inplace(st, (int64_t)M * H / 4, dHc, direct);
  linear_dgrad(st, M, G3, H, dgh, 4 * H, W[P_LD_WHH], H, dHc, H, ACC_ADD);
  if (!defer) linear_wgrad(st, M, G3, H, dgh, 4 * H, Hc, H, G[P_LD_WHH], H);
  // weight_ih gradient: x masked by the self-loop flag (XL), gathered to the active rows when compacted
  if (smode != S_ZERO) {
    const float* xl = w.XL + (size_t)vi * B * XP;
    if (x_stash) { if (!x_stashed) gather_rows(st, M, XP, rm.rows, const_cast<float*>(xl), x_stash, 0); }
    else {
      if (rm.rows) { gather_rows(st, M, XP, rm.rows, const_cast<float*>(xl), w.xc, 0); xl = w.xc; }
      linear_wgrad(st, M, G3, XP, d4, 4 * H, xl, XP, w.dWihP[1], XP);
    }
  }
  (void)Xi; (void)G;
}

// The x rows of every active (graph, step) pair of the compacted schedule, plain (xiS: combiner input) and masked by
// the self-loop flag (xlS: looper input), concatenated in step order — the operands of the deferred weight_ih gradients.
// They depend on the schedule and the true features only, so one launch fills both stashes for all 21 steps (it replaces
// two row gathers per step).  Step t belongs to node vi with vi(vi-1)/2 <= t < vi(vi+1)/2 (decode_fwd_impl's step order).
struct StepPtrV { int p[NSTEP + 1]; };
static void stash_step_x(dx_stream_t st, int B, const int32_t* step_ptr, const int* step_rows, const float* Xn,
                         const float* XL, float* xiS, float* xlS) {
  StepPtrV sp;
  for (int t = 0; t <= NSTEP; ++t) sp.p[t] = step_ptr[t];
  const int64_t nact = sp.p[NSTEP] - sp.p[0];
  if (nact <= 0) return;
  foreach (st, nact * (XP / 4), [=] DX_HD(int64_t idx) {
    const int pos = (int)(idx / (XP / 4)) + sp.p[0], c = (int)(idx % (XP / 4)) * 4;
    int t = 0;
    while (t + 1 < NSTEP && pos >= sp.p[t + 1]) ++t;
    int vi = 1;
    while ((vi + 1) * vi / 2 <= t) ++vi;
    const int64_t src = ((int64_t)vi * B + step_rows[pos]) * XP + c;
    const int64_t dst = (int64_t)(pos - sp.p[0]) * XP + c;
    st4f(xiS + dst, ld4f(Xn + src)); st4f(xlS + dst, ld4f(XL + src));
  });
}

void decode_bwd_impl(dx_stream_t st, const Weights& W, const Weights& G, int B, const float* z, const DecWs& w,
                     const Batch& bt, LossW lw) {
  const uint64_t* adj = bt.adj;
  const size_t bH = (size_t)B * H;
  // weight gradients of the re-propagates: one product per tensor over all active (graph, step) pairs at the end
  // (DX_DEFER_WGRAD=0: two products per step, for A/B)
  static const bool defer_env = [] { const char* e = getenv("DX_DEFER_WGRAD"); return !(e && e[0] == '0'); }();
  const bool defer = defer_env;
  zero_async(st, w.dHd + 6 * bH, sizeof(float) * bH);   // nothing reads the last node's state: its gradient is zero
  if (bt.step_ptr) {
    // compacted steps: msg_bwd accumulates the "out" halves for every graph, the "in" halves only on the rows of the
    // back-edge-source lists (and only those are read back): zero exactly that
    zero2d_async(st, w.dPg + H, sizeof(float) * 2 * H, sizeof(float) * H, (size_t)6 * B);
    zero2d_async(st, w.dPm + H, sizeof(float) * 2 * H, sizeof(float) * H, (size_t)6 * B);
    for (int x = 0; x < 6; ++x) {
      const int tl = NSTEP + 6 + x, n = bt.step_ptr[tl + 1] - bt.step_ptr[tl];
      if (n <= 0) continue;
      const int* rows = bt.step_rows + bt.step_ptr[tl];
      float* pg = w.dPg + (size_t)x * B * 2 * H; float* pm = w.dPm + (size_t)x * B * 2 * H;
      foreach (st, (int64_t)n * (H / 4), [=] DX_HD(int64_t i) {
        const int64_t r = rows[i / (H / 4)]; const int c = (int)(i % (H / 4)) * 4;
        st4f(pg + r * 2 * H + c, f4zero()); st4f(pm + r * 2 * H + c, f4zero());
      });
    }
  } else {
    zero_async(st, w.dPg, sizeof(float) * 6 * (size_t)B * 2 * H);
    zero_async(st, w.dPm, sizeof(float) * 6 * (size_t)B * 2 * H);
  }
  if (!bt.step_ptr) zero_async(st, w.dQ, sizeof(float) * 6 * (size_t)B * 4 * H);   // (compacted steps store dQ whole)
  for (int k = 0; k < 3; ++k) zero_async(st, w.dWihP[k], sizeof(float) * G3 * XP);
  mask_features(st, (int64_t)7 * B, B, nullptr, 0, adj, bt.Xn, w.XL);

  const bool hbatch = heads_batched(B, true);
  if (hbatch) {
    // backward of the batched parameter heads: the FIRST writer of dh_0..dh_5 (stores; every other consumer of a node
    // state accumulates on top, below)
    mlp3_bwd(st, W, G, 6 * B, P_X_W0, w.Hd, SX, w.A1[1], w.A2[1], w.dL[1], w, w.dHd, ACC_STORE);
  }
  // first propagates of nodes 1..6 as one 6B-row pass after the node loop (p1_batched; needs the in-place gate gradients)
  const bool p1b = defer && p1_batched(B, true, bt.step_ptr);
  if (p1b) zero_async(st, w.dHc06, sizeof(float) * 6 * bH);   // the second propagates (self-loop rows) add into it first
  static const bool one_stash = getenv("DX_NO_XSTASH_ONE") == nullptr;
  const bool xst = defer && bt.step_ptr && one_stash;
  if (xst) stash_step_x(st, B, bt.step_ptr, bt.step_rows, bt.Xn, w.XL, w.xiS, w.xlS);
  // small batches: the head gradients dU of the steps are kept, concatenated in step order like the states Hi, and the
  // first-layer weight gradient of the edge head (its Hi half) is ONE product over all active rows after the loop; the
  // last step of a node feeds no later head: its rows stay zero
  if (p1b) {
    const size_t na = (size_t)(bt.step_ptr[NSTEP] - bt.step_ptr[0]);
    if (na) zero_async(st, w.UCS, sizeof(float) * na * 4 * H);
  }
  int t_end = NSTEP;  // steps of node vi occupy [t_end - vi, t_end)
  for (int vi = NN - 1; vi >= 1; --vi) {
    const float* Xi = bt.Xn + (size_t)vi * B * XP;
    RowMap rm{B, B, nullptr, vi * B};
    const int t0 = t_end - vi;       // step index of vj = vi-1 ; vj = 0 is t_end-1
    float* const dHi = w.dHd + (size_t)vi * bH;   // the node's state gradient is consumed in place (no later reader of dHd[vi])
    if (!bt.step_ptr) zero_async(st, w.dHrun, sizeof(float) * bH);   // (compacted steps accumulate first-touch, below)
    const bool compact = bt.step_ptr != nullptr;
    if (compact) {
      // dHi is the running gradient of the node's current state; a step consumes and clears it on its
      // active rows.  The gradient of U (the state's edge-head product) is never materialised for all
      // graphs: the version of U written by step (vi,vj) is read by the heads (vi,vj-1), (vi,vj-2), ... up to
      // and including the graph's next active step, and head_sum re-forms exactly that sum for the active rows.
      auto head_list = [&](HeadSumP& hs, int vj_first) {     // heads (vi, vj_first), (vi, vj_first-1), ..., (vi, 0)
        hs.n = 0;
        for (int x = vj_first; x >= 0; --x, ++hs.n) {
          const int tx = t0 + (vi - 1 - x);
          hs.mask[hs.n] = reinterpret_cast<const uint8_t*>(w.E1[tx]); hs.dl[hs.n] = w.dl2[tx]; hs.vj[hs.n] = x;
        }
      };
      for (int vj = 0; vj < vi; ++vj) {
        const int t = t0 + (vi - 1 - vj);
        const int n = bt.step_ptr[t + 1] - bt.step_ptr[t];
        if (n <= 0) continue;
        const int* rows = bt.step_rows + bt.step_ptr[t];
        RowMap rc{n, B, rows, vi * B};
        gather_rows(st, n, H, rows, dHi, w.dHiC, 1);
        if (vj > 0) {                                        // (the last step's state feeds no later head)
          float* const dUs = p1b ? w.UCS + (size_t)(bt.step_ptr[t] - bt.step_ptr[0]) * 4 * H : w.UC;
          HeadSumP hs{n, rows, adj, vi, 0, 1, {}, {}, {}, W[P_E_W2], dUs};
          head_list(hs, vj - 1);
          head_sum(st, hs);
          linear_dgrad(st, n, 4 * H, H, dUs, 4 * H, W[P_E_W0], 2 * H, w.dHiC, H, ACC_ADD);
          if (!p1b) linear_wgrad(st, n, 4 * H, H, dUs, 4 * H, w.Hi[t], H, G[P_E_W0], 2 * H);
        }
        const size_t soff = (size_t)(bt.step_ptr[t] - bt.step_ptr[0]) * XP;   // this step's rows in the concatenated x stashes
        if (defer) {
        looper_bwd(st, W, G, B, vi, rc, w.dHiC, w.g_l[t], w.Hc[t], S_SELF, adj, Xi, w, w.dHc, false, w.xlS + soff, false, xst);
        // combiner: gate gradients in place over g_c[t]; its two weight gradients are deferred like the looper's
        CellBwd cc{rc, w.dHc, 0, w.g_c[t], 0, w.Hin[t], 0, w.g_c[t], nullptr, w.g_c[t] + H, w.dHin, S_ONE, adj};
        cell_bwd(st, cc, G[P_CD_BIH], G[P_CD_BHH]);
        linear_dgrad(st, n, G3, H, w.g_c[t] + H, 4 * H, W[P_CD_WHH], H, w.dHin, H, ACC_ADD);
        if (!xst) gather_rows(st, n, XP, rows, const_cast<float*>(Xi), w.xiS + soff, 0);
        } else {
        looper_bwd(st, W, G, B, vi, rc, w.dHiC, w.g_l[t], w.Hc[t], S_SELF, adj, Xi, w, w.dHc, false);
        CellBwd cc{rc, w.dHc, 0, w.g_c[t], 0, w.Hin[t], 0, w.dgx, nullptr, w.dgh, w.dHin, S_ONE, adj};
        cell_bwd(st, cc, G[P_CD_BIH], G[P_CD_BHH]);
        linear_dgrad(st, n, G3, H, w.dgh, 4 * H, W[P_CD_WHH], H, w.dHin, H, ACC_ADD);
        linear_wgrad(st, n, G3, H, w.dgh, 4 * H, w.Hin[t], H, G[P_CD_WHH], H);
        gather_rows(st, n, XP, rows, const_cast<float*>(Xi), w.xc, 0);
        linear_wgrad(st, n, G3, XP, w.dgx, 4 * H, w.xc, XP, w.dWihP[0], XP);
        }
        {
          // dHrun[b] += dHin[m]: the running gradient of the aggregate.  Steps are walked with vj ascending, so a row was
          // written before iff the graph has an edge between vi and some node below vj: first touch stores (no zero fill)
          const float* src = w.dHin; float* dst = w.dHrun;
          foreach (st, (int64_t)n * (H / 4), [=] DX_HD(int64_t idx) {
            const int m = (int)(idx / (H / 4)), c = (int)(idx % (H / 4)) * 4;
            const int64_t b = rows[m];
            const uint64_t A = adj[b];
            bool add = false;
            for (int x = 0; x < vj; ++x) add = add || (abit(A, x, vi) | abit(A, vi, x));
            float4 v = ld4f(src + (int64_t)m * H + c);
            if (add) { const float4 o = ld4f(dst + b * H + c); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            st4f(dst + b * H + c, v);
          });
        }
        RowMap rs{n, B, rows, vj * B};
        MsgBwd mb{rs, w.Pg, w.Pm, W[P_G_B], adj, w.dHrun, 0, w.dPg, w.dPm, G[P_G_B], 1, vi, vi, 1};
        mb.lazy_in = 1;                                          // the "in" half only exists on back-edge rows
        msg_bwd(st, mb);
      }
      // U = Hi_p2 W^T (the state every graph had before its first edge): read by the heads up to the first active step
      float* dU = p1b ? w.dU6 + (size_t)(vi - 1) * B * 4 * H : w.dE1;
      HeadSumP hs{B, nullptr, adj, vi, 0, 1, {}, {}, {}, W[P_E_W2], dU};
      head_list(hs, vi - 1);
      head_sum(st, hs);
      if (!p1b) {                                                // (small batches: both products once over 6B rows after the loop)
      linear_wgrad(st, B, 4 * H, H, dU, 4 * H, w.Hi_p2[vi], H, G[P_E_W0], 2 * H);
      linear_dgrad(st, B, 4 * H, H, dU, 4 * H, W[P_E_W0], 2 * H, dHi, H, ACC_ADD);
      }
      // dQ of node j = vi-1 (consumed below): heads (vi', j) of every later node vi' (all of them are done)
      {
        const int j = vi - 1;
        HeadSumP hq{B, nullptr, adj, vi, 0, 0, {}, {}, {}, W[P_E_W2], w.dQ + (size_t)j * B * 4 * H};
        for (int v2 = NN - 1; v2 > j; --v2, ++hq.n) {
          const int tx = v2 * (v2 - 1) / 2 + (v2 - 1 - j);
          hq.mask[hq.n] = reinterpret_cast<const uint8_t*>(w.E1[tx]); hq.dl[hq.n] = w.dl2[tx]; hq.vj[hq.n] = j;
        }
        head_sum(st, hq);
      }
    } else
    for (int vj = 0; vj < vi; ++vj) {
      const int t = t0 + (vi - 1 - vj);
      // looper then combiner of this propagate
      looper_bwd(st, W, G, B, vi, rm, dHi, w.g_l[t], w.Hc[t], S_SELF, adj, Xi, w, w.dHc, false);
      CellBwd cc{rm, w.dHc, 0, w.g_c[t], 0, w.Hin[t], 0, w.dgx, nullptr, w.dgh, w.dHin, S_ONE, adj};
      cell_bwd(st, cc, G[P_CD_BIH], G[P_CD_BHH]);
      linear_dgrad(st, B, G3, H, w.dgh, 4 * H, W[P_CD_WHH], H, w.dHin, H, ACC_ADD);
      linear_wgrad(st, B, G3, H, w.dgh, 4 * H, w.Hin[t], H, G[P_CD_WHH], H);
      linear_wgrad(st, B, G3, XP, w.dgx, 4 * H, Xi, XP, w.dWihP[0], XP);
      add_inplace(st, (int64_t)bH / 4, w.dHrun, w.dHin);
      // message of vj was part of this and every later aggregate of node vi
      RowMap rs{B, B, nullptr, vj * B};
      MsgBwd mb{rs, w.Pg, w.Pm, W[P_G_B], adj, w.dHrun, 0, w.dPg, w.dPm, G[P_G_B], 1, vi, vi, 1};
      msg_bwd(st, mb);
      // edge head of this step read the PREVIOUS Hi (step t-1, or P2 when vj = vi-1)
      const float* Hi_prev = (vj == vi - 1) ? w.Hi_p2[vi] : w.Hi[t - 1];
      linear_wgrad(st, B, 2, 4 * H, w.dl2[t], LD_E, w.E1[t], 4 * H, G[P_E_W2], 4 * H);
      colsum_accum(st, B, 2, w.dl2[t], LD_E, G[P_E_B2]);
      relu_head_bwd(st, B, 4 * H, 2, w.E1[t], w.dl2[t], LD_E, W[P_E_W2], w.dE1, w.dQ + (size_t)vj * B * 4 * H);
      linear_wgrad(st, B, 4 * H, H, w.dE1, 4 * H, Hi_prev, H, G[P_E_W0], 2 * H);
      linear_dgrad(st, B, 4 * H, H, w.dE1, 4 * H, W[P_E_W0], 2 * H, dHi, H, ACC_STORE);
    }
    // dHi now holds the gradient of Hi_p2.  P2 and P1 share Hc0.
    int ns = 0; const int* rows_s = nullptr;
    if (p1b) {
      // small batches: dHd[vi] is left as it is (the gradient of Hi_p2 but for the product U = Hi_p2 W^T, added for all six
      // nodes at once after the loop); the second and first propagates of all nodes follow there
    } else if (compact) {
      // P2 ran on the self-loop rows only: their gradient goes through P2's looper (compact), every other
      // row's gradient passes straight to Hi_p1 (Hi_p2 == Hi_p1 there) and joins the self-loop head's.
      const int ts = NSTEP + vi - 1;
      ns = bt.step_ptr[ts + 1] - bt.step_ptr[ts]; rows_s = bt.step_rows + bt.step_ptr[ts];
      if (ns > 0) {
        RowMap rs{ns, B, rows_s, vi * B};
        gather_rows(st, ns, H, rows_s, dHi, w.dHiC, 1);
        gather_rows(st, ns, H, rows_s, w.Hc0[vi], w.Hrun, 0);          // (Hrun is free scratch in the backward pass)
        looper_bwd(st, W, G, B, vi, rs, w.dHiC, w.g_p2[vi], w.Hrun, S_SELF, adj, Xi, w, w.dHc, false);
      }
    } else {
      looper_bwd(st, W, G, B, vi, rm, dHi, w.g_p2[vi], w.Hc0[vi], S_SELF, adj, Xi, w, w.dHc0, false);
    }
    if (!p1b) {
    // self-loop head consumed Hi_p1
    linear_wgrad(st, B, 1, 2 * H, w.dls[vi], LD_E, w.ES1[vi], 2 * H, G[P_ES_W2], 2 * H);
    colsum_accum(st, B, 1, w.dls[vi], LD_E, G[P_ES_B2]);
    relu_head_bwd(st, B, 2 * H, 1, w.ES1[vi], w.dls[vi], LD_E, W[P_ES_W2], w.dES1, nullptr);
    linear_wgrad(st, B, 2 * H, H, w.dES1, 2 * H, w.Hi_p1[vi], H, G[P_ES_W0], H);
    colsum_accum(st, B, 2 * H, w.dES1, 2 * H, G[P_ES_B0]);
    linear_dgrad(st, B, 2 * H, H, w.dES1, 2 * H, W[P_ES_W0], H, dHi, H, compact ? ACC_ADD : ACC_STORE);
    // first propagate of the node (all B graphs): gate gradients in place over g_p1[vi] / g_c0[vi]; the per-node buffers
    // of nodes 1..6 are contiguous, so both weight gradients are formed once over 6B rows after the node loop
    looper_bwd(st, W, G, B, vi, rm, dHi, w.g_p1[vi], w.Hc0[vi], S_ZERO, adj, Xi, w, w.dHc0, !compact, nullptr, defer);
    if (ns > 0) scatter_rows(st, ns, H, rows_s, w.dHc, w.dHc0, 1);
    // combiner with H_in = 0: only input weights / biases receive gradient
    float* d0 = defer ? w.g_c0[vi] : w.dgx;
    CellBwd c0{rm, w.dHc0, 0, w.g_c0[vi], 0, nullptr, 0, d0, nullptr, d0 + H, nullptr, S_ONE, adj};
    cell_bwd(st, c0, G[P_CD_BIH], G[P_CD_BHH]);
    if (!defer) linear_wgrad(st, B, G3, XP, w.dgx, 4 * H, Xi, XP, w.dWihP[0], XP);
    }
    // parameter head of node vi read h_{vi-1}
    float* dprev = w.dHd + (size_t)(vi - 1) * bH;
    const float* hprev = w.Hd + (size_t)(vi - 1) * bH;
    // first writer of dh_{vi-1} (its other consumers are folded in just below): stores, so dHd needs no zero fill
    if (!hbatch) mlp3_bwd(st, W, G, B, P_X_W0, hprev, SX, w.A1[vi], w.A2[vi], w.dL[vi], w, dprev, ACC_STORE);
    // every consumer of node vi-1 is done: fold its projection gradients into dh_{vi-1}
    const int j = vi - 1;
    const float* dPg = w.dPg + (size_t)j * B * 2 * H; const float* dPm = w.dPm + (size_t)j * B * 2 * H;
    const float* dQ = w.dQ + (size_t)j * B * 4 * H;
    for (int half = 0; half < 2; ++half) {
      if (half == HALF_IN && compact) {
        // the "in" half was only produced (and only received gradient) on the back-edge-source rows of node j
        const int tl = NSTEP + 6 + j, n = bt.step_ptr[tl + 1] - bt.step_ptr[tl];
        if (n <= 0) continue;
        const int* rows = bt.step_rows + bt.step_ptr[tl];
        float* tg = w.UC; float* tm = w.UC + bH; float* th = w.UC + 2 * bH; float* tdx = w.UC + 3 * bH;
        gather_rows(st, n, H, rows, const_cast<float*>(dPg), tg, 0, 2 * H);
        gather_rows(st, n, H, rows, const_cast<float*>(dPm), tm, 0, 2 * H);
        gather_rows(st, n, H, rows, const_cast<float*>(hprev), th, 0);
        linear_dgrad(st, n, H, H, tg, H, W[P_G_W], 2 * H, tdx, H, ACC_STORE);
        linear_dgrad(st, n, H, H, tm, H, W[P_M_W], 2 * H, tdx, H, ACC_ADD);
        scatter_rows(st, n, H, rows, tdx, dprev, 1);
        linear_wgrad(st, n, H, H, tg, H, th, H, G[P_G_W], 2 * H);
        linear_wgrad(st, n, H, H, tm, H, th, H, G[P_M_W], 2 * H);
        continue;
      }
      proj_dgrad(st, B, dPg, W[P_G_W], dprev, half, ACC_ADD);
      proj_dgrad(st, B, dPm, W[P_M_W], dprev, half, ACC_ADD);
      if (p1b) continue;                                          // (weight gradients: once over 6B rows after the loop)
      proj_wgrad(st, B, dPg, hprev, G[P_G_W], half);
      proj_wgrad(st, B, dPm, hprev, G[P_M_W], half);
    }
    linear_dgrad(st, B, 4 * H, H, dQ, 4 * H, W[P_E_W0] + H, 2 * H, dprev, H, ACC_ADD);
    if (!p1b) linear_wgrad(st, B, 4 * H, H, dQ, 4 * H, hprev, H, G[P_E_W0] + H, 2 * H);
    if (!compact) colsum_accum(st, B, 4 * H, dQ, 4 * H, G[P_E_B0]);   // (compacted steps: the fused forward head summed it)
    t_end = t0;
  }
  // root cell: h_0 = GRU_root(x0, H_init)
  {
    RowMap rm{B, B, nullptr, 0};
    CellBwd cr{rm, w.dHd, 0, w.g_root, 0, w.Hinit, 0, w.dgx, nullptr, w.dgh, w.dHinit, S_ONE, adj};
    cell_bwd(st, cr, G[P_RD_BIH], G[P_RD_BHH]);
    linear_dgrad(st, B, G3, H, w.dgh, 4 * H, W[P_RD_WHH], H, w.dHinit, H, ACC_ADD);
    linear_wgrad(st, B, G3, H, w.dgh, 4 * H, w.Hinit, H, G[P_RD_WHH], H);
    linear_wgrad(st, B, G3, XP, w.dgx, 4 * H, bt.Xn, XP, w.dWihP[2], XP);
  }
  mlp3_bwd(st, W, G, B, P_X0_W0, w.Hinit, SX0 + 32, w.A1[0], w.A2[0], w.dL[0], w, w.dHinit);
  // tanh' from the recomputed pre-activation (Hrun is free scratch here; same kernel and operands as the forward pass)
  linear_fwd(st, B, H, Z, z, Z, W[P_ZH_W], Z, W[P_ZH_B], w.Hrun, H);
  tanh_bwd_pre(st, (int64_t)bH, w.dHinit, w.Hrun);
  linear_wgrad(st, B, H, Z, w.dHinit, H, z, Z, G[P_ZH_W], Z);
  colsum_accum(st, B, H, w.dHinit, H, G[P_ZH_B]);
  linear_dgrad(st, B, H, Z, w.dHinit, H, W[P_ZH_W], Z, w.dz, Z, ACC_STORE);
  if (p1b) {
    // self-loop heads, looper and combiner of the first propagates of nodes 1..6: one pass over 6B rows (rows [B, 7B) of
    // dHd hold the gradients of Hi_p2; every per-node buffer below is carved back to back)
    const int B6 = 6 * B;
    float* dH6 = w.dHd + bH;
    RowMap rm6{B6, B, nullptr, B};
    // U = Hi_p2 W^T of every node (the state each graph had before its first edge): head gradients kept in dU6
    linear_wgrad(st, B6, 4 * H, H, w.dU6, 4 * H, w.Hi_p2[1], H, G[P_E_W0], 2 * H);
    linear_dgrad(st, B6, 4 * H, H, w.dU6, 4 * H, W[P_E_W0], 2 * H, dH6, H, ACC_ADD);
    // second propagates (self-loop rows of each node, compact): their gradient goes through P2's looper, every other row's
    // passes straight to Hi_p1 (Hi_p2 == Hi_p1 there); the combiner-state gradient waits in dHc06[vi]
    for (int vi = NN - 1; vi >= 1; --vi) {
      const int ts = NSTEP + vi - 1, ns = bt.step_ptr[ts + 1] - bt.step_ptr[ts];
      if (ns <= 0) continue;
      const int* rows_s = bt.step_rows + bt.step_ptr[ts];
      RowMap rs{ns, B, rows_s, vi * B};
      gather_rows(st, ns, H, rows_s, w.dHd + (size_t)vi * bH, w.dHiC, 1);
      gather_rows(st, ns, H, rows_s, w.Hc0[vi], w.Hrun, 0);            // (Hrun is free scratch in the backward pass)
      looper_bwd(st, W, G, B, vi, rs, w.dHiC, w.g_p2[vi], w.Hrun, S_SELF, adj, bt.Xn + (size_t)vi * B * XP, w, w.dHc, false);
      scatter_rows(st, ns, H, rows_s, w.dHc, w.dHc06 + (size_t)(vi - 1) * bH, 1);
    }
    linear_wgrad(st, B6, 1, 2 * H, w.dls[1], LD_E, w.ES1[1], 2 * H, G[P_ES_W2], 2 * H);
    colsum_accum(st, B6, 1, w.dls[1], LD_E, G[P_ES_B2]);
    relu_head_bwd(st, B6, 2 * H, 1, w.ES1[1], w.dls[1], LD_E, W[P_ES_W2], w.dES16, nullptr);
    linear_wgrad(st, B6, 2 * H, H, w.dES16, 2 * H, w.Hi_p1[1], H, G[P_ES_W0], H);
    colsum_accum(st, B6, 2 * H, w.dES16, 2 * H, G[P_ES_B0]);
    linear_dgrad(st, B6, 2 * H, H, w.dES16, 2 * H, W[P_ES_W0], H, dH6, H, ACC_ADD);
    // looper (x_loop = 0): gate gradients in place over g_p1, dHc0 += dHi * z + dgh W_hh
    CellBwd cb{rm6, dH6, 0, w.g_p1[1], 0, w.Hc0[1], 0, w.g_p1[1], nullptr, w.g_p1[1] + H, w.dir6, S_ZERO, adj};
    cell_bwd(st, cb, G[P_LD_BIH], G[P_LD_BHH]);
    add_inplace(st, (int64_t)B6 * H / 4, w.dHc06, w.dir6);
    linear_dgrad(st, B6, G3, H, w.g_p1[1] + H, 4 * H, W[P_LD_WHH], H, w.dHc06, H, ACC_ADD);
    // combiner with H_in = 0: only input weights / biases receive gradient
    CellBwd c0{rm6, w.dHc06, 0, w.g_c0[1], 0, nullptr, 0, w.g_c0[1], nullptr, w.g_c0[1] + H, nullptr, S_ONE, adj};
    cell_bwd(st, c0, G[P_CD_BIH], G[P_CD_BHH]);
    // weight gradients of what reads the finished node states h_0..h_5 (the "out" halves of the gate / mapper projections
    // and the Hj half of the edge head's first layer): dPg / dPm / dQ and Hd are node-major, one product each over 6B rows
    proj_wgrad(st, B6, w.dPg, w.Hd, G[P_G_W], HALF_OUT);
    proj_wgrad(st, B6, w.dPm, w.Hd, G[P_M_W], HALF_OUT);
    linear_wgrad(st, B6, 4 * H, H, w.dQ, 4 * H, w.Hd, H, G[P_E_W0] + H, 2 * H);
  }
  if (defer) {
    // first propagates of nodes 1..6 (g_p1 / g_c0 / Hc0 are per-node buffers carved back to back; Xn is node-major)
    linear_wgrad(st, 6 * B, G3, H, w.g_p1[1] + H, 4 * H, w.Hc0[1], H, G[P_LD_WHH], H);
    linear_wgrad(st, 6 * B, G3, XP, w.g_c0[1], 4 * H, bt.Xn + (size_t)B * XP, XP, w.dWihP[0], XP);
  }
  if (bt.step_ptr && defer) {
    // the deferred weight gradients of the 21 re-propagates: ONE product per tensor over every active (graph, step) pair
    // (the per-step gate-gradient, state and x buffers are contiguous in step order, see carve_dec)
    const int nact = bt.step_ptr[NSTEP] - bt.step_ptr[0];
    if (nact > 0) {
      linear_wgrad(st, nact, G3, H, w.g_l[0] + H, 4 * H, w.Hc[0], H, G[P_LD_WHH], H);
      linear_wgrad(st, nact, G3, XP, w.g_l[0], 4 * H, w.xlS, XP, w.dWihP[1], XP);
      linear_wgrad(st, nact, G3, H, w.g_c[0] + H, 4 * H, w.Hin[0], H, G[P_CD_WHH], H);
      linear_wgrad(st, nact, G3, XP, w.g_c[0], 4 * H, w.xiS, XP, w.dWihP[0], XP);
      if (p1b) linear_wgrad(st, nact, 4 * H, H, w.UCS, 4 * H, w.Hi[0], H, G[P_E_W0], 2 * H);
    }
  }
  unpad_add_wih(st, w.dWihP[0], SX, G[P_CD_WIH]);
  unpad_add_wih(st, w.dWihP[1], SX, G[P_LD_WIH]);
  unpad_add_wih(st, w.dWihP[2], SX0, G[P_RD_WIH]);
  (void)lw;
}

}  // namespace dx
