// Level-scheduled encoder: model.py:200-212 (encode) / :151-198 (_propagate, encode=True).
//
// The reference walks v = 6..0 over the whole batch and asks every DGL graph for its
// neighbours from Python.  Here the batcher's level schedule groups operator rows
// (v*B+b) by topological level; each level is ONE fused step over its row list:
//   gated-sum aggregation (feedback back-edges arrive through the "out" half of the
//   projections) -> combiner GRU -> looper GRU (input masked by the self-loop flag)
//   -> gate/mapper projections of the new state, reused by every lower neighbour.
// The root step (node 0 of all graphs) runs last, then the two latent heads.
// Activations live in schedule order (level-major), so every product is a dense row range:
// no gathers in the GEMMs, which lets them run on the TMA/tcgen05 path.
#include "dx_engine.h"

namespace dx {

// level_ptr (HOST, n_levels + 1; optional): the per-level temporaries (input / hidden gate products, gate gradients) are
// then sized for the largest level of THIS schedule (and the B rows of the root step) instead of the 6B-row worst case.
EncWs carve_enc(Arena& ar, int64_t B, bool train, int n_levels, const int32_t* level_ptr) {
  EncWs w{};
  const size_t R7 = (size_t)7 * B, R6 = (size_t)6 * B;
  size_t RT = R6;                                                   // rows of a per-level temporary
  if (level_ptr && n_levels > 0) {
    RT = (size_t)B;
    for (int L = 0; L < n_levels; ++L) { const size_t m = (size_t)(level_ptr[L + 1] - level_ptr[L]); if (m > RT) RT = m; }
  }
  w.Hin = ar.take<float>(R7 * H); w.Hc = ar.take<float>(R7 * H); w.Hv = ar.take<float>(R7 * H);
  w.Pg = ar.take<float>(R7 * 2 * H); w.Pm = ar.take<float>(R7 * 2 * H);
  w.gxc = ar.take<float>(RT * G3); w.gxl = ar.take<float>(RT * G3); w.gh = ar.take<float>(RT * G3);
  w.XnS = ar.take<float>(R6 * XP); w.pos = ar.take<int>(R7);
  for (int k = 0; k < 3; ++k) w.WihP[k] = ar.take<float>((size_t)G3 * XP);
  if (train) {
    for (int k = 0; k < 3; ++k) w.dWihP[k] = ar.take<float>((size_t)G3 * XP);
    w.XnSL = ar.take<float>(R6 * XP);
    w.gc = ar.take<float>(R7 * 4 * H); w.gl = ar.take<float>(R7 * 4 * H);
    w.dH = ar.take<float>(R7 * H); w.dHin = ar.take<float>(R7 * H);
    w.dPg = ar.take<float>(R6 * 2 * H); w.dPm = ar.take<float>(R6 * 2 * H); w.dgb = nullptr;
    w.dgx = ar.take<float>(RT * 4 * H); w.dgh = w.dgx + H; w.dgxs = nullptr;   // one D4 buffer, two views (CellBwd)
    w.dHc = ar.take<float>(RT * H); w.dsraw = ar.take<float>((size_t)B * Z);
  }
  return w;
}

void encode_fwd_impl(dx_stream_t st, const Weights& W, const Batch& bt, const EncWs& w, float* mu, float* sd,
                     bool train) {
  const int B = (int)bt.B;
  const int64_t R6 = (int64_t)6 * B;
  // schedule-order bookkeeping: inverse map and the permuted feature rows
  {
    int* pos = w.pos; const int32_t* rows = bt.level_rows; const float* Xn = bt.Xn; float* XnS = w.XnS;
    foreach (st, (int64_t)7 * B, [=] DX_HD(int64_t i) {
      if (i < R6) pos[rows[i]] = (int)i; else pos[i - R6] = (int)i;     // node 0 of graph b sits at 6B+b
    });
    foreach (st, R6 * (XP / 4), [=] DX_HD(int64_t i) {
      const int64_t p = i / (XP / 4); const int c = (int)(i % (XP / 4)) * 4;
      st4f(XnS + p * XP + c, ld4f(Xn + (int64_t)rows[p] * XP + c));
    });
  }
  // input weights: training and the tensor-core modes use 32-column padded copies (TMA-addressable, exact: the
  // padded columns are zero); FP32 FFMA inference reads the blob
  const float* Wc = W[P_CE_WIH]; const float* Wl = W[P_LE_WIH]; const float* Wr = W[P_RE_WIH];
  int Kx = SX, Kr = SX0, ldx = SX, ldr = SX0;
  if (train || get_precision() != PREC_FP32) {
    pad_wih(st, W[P_CE_WIH], SX, w.WihP[0]); pad_wih(st, W[P_LE_WIH], SX, w.WihP[1]); pad_wih(st, W[P_RE_WIH], SX0, w.WihP[2]);
    Wc = w.WihP[0]; Wl = w.WihP[1]; Wr = w.WihP[2]; Kx = Kr = ldx = ldr = XP;
  }
  for (int L = 0; L < bt.n_levels; ++L) {
    const int base = bt.level_ptr[L];
    const int M = bt.level_ptr[L + 1] - base;
    if (M <= 0) continue;
    RowMap rm{M, B, bt.level_rows + base, 0};
    float* Hin = w.Hin + (size_t)base * H; float* Hc = w.Hc + (size_t)base * H; float* Hv = w.Hv + (size_t)base * H;
    if (L > 0) {
      MsgFwd mf{rm, w.Pg, w.Pm, W[P_G_B], bt.adj, Hin, 0, -1, 0, 0};
      mf.pos = w.pos;
      msg_fwd(st, mf);
    }
    linear_fwd(st, M, G3, Kx, w.XnS + (size_t)base * XP, XP, Wc, ldx, nullptr, w.gxc, G3);
    linear_fwd(st, M, G3, Kx, w.XnS + (size_t)base * XP, XP, Wl, ldx, nullptr, w.gxl, G3);
    if (L > 0) linear_fwd(st, M, G3, H, Hin, H, W[P_CE_WHH], H, nullptr, w.gh, G3);
    CellFwd c1{rm, w.gxc, L > 0 ? w.gh : nullptr, W[P_CE_BIH], W[P_CE_BHH], L > 0 ? Hin : nullptr, 0, Hc, 0,
               train ? w.gc + (size_t)base * 4 * H : nullptr, 0, S_ONE, bt.adj};
    cell_fwd(st, c1);
    linear_fwd(st, M, G3, H, Hc, H, W[P_LE_WHH], H, nullptr, w.gh, G3);
    CellFwd c2{rm, w.gxl, w.gh, W[P_LE_BIH], W[P_LE_BHH], Hc, 0, Hv, 0, train ? w.gl + (size_t)base * 4 * H : nullptr, 0,
               S_SELF, bt.adj};
    cell_fwd(st, c2);
    // projections of the finished state, by half (see proj_fwd): a lower node v reads the "in" half of x through
    // its forward edge x -> v and the "out" half only through a feedback back-edge v -> x, so the "out" half is
    // computed on the level's leading rows alone (the batcher puts the back-edge targets first)
    const int Mr = bt.level_rare ? bt.level_rare[L] : M;
    proj_fwd(st, M, Hv, W[P_G_W], w.Pg + (size_t)base * 2 * H, HALF_IN);
    proj_fwd(st, M, Hv, W[P_M_W], w.Pm + (size_t)base * 2 * H, HALF_IN);
    if (Mr > 0) {
      proj_fwd(st, Mr, Hv, W[P_G_W], w.Pg + (size_t)base * 2 * H, HALF_OUT);
      proj_fwd(st, Mr, Hv, W[P_M_W], w.Pm + (size_t)base * 2 * H, HALF_OUT);
    }
  }
  // root step: node 0 of every graph (positions 6B..7B-1)
  RowMap r0{B, B, nullptr, 0};
  float* Hin0 = w.Hin + (size_t)R6 * H; float* Hv0 = w.Hv + (size_t)R6 * H;
  MsgFwd mf{r0, w.Pg, w.Pm, W[P_G_B], bt.adj, Hin0, 0, -1, 0, 0};
  mf.pos = w.pos;
  msg_fwd(st, mf);
  linear_fwd(st, B, G3, Kr, bt.Xn, XP, Wr, ldr, nullptr, w.gxc, G3);
  linear_fwd(st, B, G3, H, Hin0, H, W[P_RE_WHH], H, nullptr, w.gh, G3);
  CellFwd cr{r0, w.gxc, w.gh, W[P_RE_BIH], W[P_RE_BHH], Hin0, 0, Hv0, 0, train ? w.gc + (size_t)R6 * 4 * H : nullptr, 0,
             S_ONE, bt.adj};
  cell_fwd(st, cr);
  linear_fwd(st, B, Z, H, Hv0, H, W[P_MU_W], H, W[P_MU_B], mu, Z);
  linear_fwd(st, B, Z, H, Hv0, H, W[P_STD_W], H, W[P_STD_B], sd, Z, ACT_SOFTPLUS);
}

// Backward of the above.  dmu, dstd: (B,128).  Accumulates into the gradient blob G.
void encode_bwd_impl(dx_stream_t st, const Weights& W, const Weights& G, const Batch& bt, const EncWs& w,
                     const float* dmu, const float* dstd, const float* sd) {
  const int B = (int)bt.B;
  const int64_t R6 = (int64_t)6 * B;
  float* Hin0 = w.Hin + (size_t)R6 * H; float* Hv0 = w.Hv + (size_t)R6 * H;
  float* dH0 = w.dH + (size_t)R6 * H; float* dHin0 = w.dHin + (size_t)R6 * H;
  for (int k = 0; k < 3; ++k) zero_async(st, w.dWihP[k], sizeof(float) * G3 * XP);
  mask_features(st, R6, B, bt.level_rows, 0, bt.adj, w.XnS, w.XnSL);
  // softplus': sigmoid(raw) = 1 - exp(-std)
  {
    float* dsraw = w.dsraw;
    foreach (st, (int64_t)B * Z, [=] DX_HD(int64_t i) { dsraw[i] = dstd[i] * (1.f - expf(-sd[i])); });
  }
  linear_dgrad(st, B, Z, H, dmu, Z, W[P_MU_W], H, dH0, H, ACC_STORE);
  linear_dgrad(st, B, Z, H, w.dsraw, Z, W[P_STD_W], H, dH0, H, ACC_ADD);
  linear_wgrad(st, B, Z, H, dmu, Z, Hv0, H, G[P_MU_W], H);
  linear_wgrad(st, B, Z, H, w.dsraw, Z, Hv0, H, G[P_STD_W], H);
  colsum_accum(st, B, Z, dmu, Z, G[P_MU_B]);
  colsum_accum(st, B, Z, w.dsraw, Z, G[P_STD_B]);
  // root cell
  RowMap r0{B, B, nullptr, 0};
  CellBwd cr{r0, dH0, 0, w.gc + (size_t)R6 * 4 * H, 0, Hin0, 0, w.dgx, nullptr, w.dgh, dHin0, S_ONE, bt.adj};
  cell_bwd(st, cr, G[P_RE_BIH], G[P_RE_BHH]);
  linear_dgrad(st, B, G3, H, w.dgh, 4 * H, W[P_RE_WHH], H, dHin0, H, ACC_ADD);
  linear_wgrad(st, B, G3, H, w.dgh, 4 * H, Hin0, H, G[P_RE_WHH], H);
  linear_wgrad(st, B, G3, XP, w.dgx, 4 * H, bt.Xn, XP, w.dWihP[2], XP);

  for (int L = bt.n_levels - 1; L >= 0; --L) {
    const int base = bt.level_ptr[L];
    const int M = bt.level_ptr[L + 1] - base;
    if (M <= 0) continue;
    RowMap rm{M, B, bt.level_rows + base, 0};
    float* Hin = w.Hin + (size_t)base * H; float* Hc = w.Hc + (size_t)base * H; float* Hv = w.Hv + (size_t)base * H;
    float* dH = w.dH + (size_t)base * H; float* dHin = w.dHin + (size_t)base * H;
    const float* Xs = w.XnS + (size_t)base * XP;
    // gradients reaching the projections of these source rows from every lower neighbour
    MsgBwd mb{rm, w.Pg + (size_t)base * 2 * H, w.Pm + (size_t)base * 2 * H, W[P_G_B], bt.adj, w.dHin, 0, w.dPg, w.dPm,
              G[P_G_B], 0, -1, 0, 0};
    mb.pos = w.pos; mb.p_compact = 1;
    mb.out_rows = bt.level_rare ? bt.level_rare[L] : -1;      // rows past the back-edge-target prefix never read their "out" half
    msg_bwd(st, mb);
    const int Mr = bt.level_rare ? bt.level_rare[L] : M;      // rows whose "out" half exists (leading rows of the level)
    for (int half = 0; half < 2; ++half) {
      const int Mh = half == HALF_IN ? M : Mr;
      if (Mh <= 0) continue;
      proj_dgrad(st, Mh, w.dPg, W[P_G_W], dH, half, half == HALF_IN ? ACC_STORE : ACC_ADD);
      proj_dgrad(st, Mh, w.dPm, W[P_M_W], dH, half, ACC_ADD);
      proj_wgrad(st, Mh, w.dPg, Hv, G[P_G_W], half);
      proj_wgrad(st, Mh, w.dPm, Hv, G[P_M_W], half);
    }
    // looper: the gate gradients are written IN PLACE over the level's saved gates (schedule order: the levels are one
    // contiguous matrix), the cells' weight gradients are formed once over all levels after the loop
    float* gl = w.gl + (size_t)base * 4 * H;
    CellBwd cl{rm, dH, 0, gl, 0, Hc, 0, gl, nullptr, gl + H, w.dHc, S_SELF, bt.adj};
    cell_bwd(st, cl, G[P_LE_BIH], G[P_LE_BHH]);
    linear_dgrad(st, M, G3, H, gl + H, 4 * H, W[P_LE_WHH], H, w.dHc, H, ACC_ADD);
    // combiner
    float* gc = w.gc + (size_t)base * 4 * H;
    CellBwd cc{rm, w.dHc, 0, gc, 0, L > 0 ? Hin : nullptr, 0, gc, nullptr, gc + H, L > 0 ? dHin : nullptr, S_ONE, bt.adj};
    cell_bwd(st, cc, G[P_CE_BIH], G[P_CE_BHH]);
    if (L > 0) linear_dgrad(st, M, G3, H, gc + H, 4 * H, W[P_CE_WHH], H, dHin, H, ACC_ADD);
    (void)Xs;
  }
  {
    // one weight-gradient product per tensor over the 6B operator rows (x masked by the self-loop flag for the looper;
    // the combiner's hidden product exists from level 1 on: level-0 rows have H_in = 0 and no saved input state)
    const int n_op = (int)R6;
    const int b1 = bt.n_levels > 1 ? bt.level_ptr[1] : n_op;
    linear_wgrad(st, n_op, G3, H, w.gl + H, 4 * H, w.Hc, H, G[P_LE_WHH], H);
    linear_wgrad(st, n_op, G3, XP, w.gl, 4 * H, w.XnSL, XP, w.dWihP[1], XP);
    if (n_op > b1) linear_wgrad(st, n_op - b1, G3, H, w.gc + (size_t)b1 * 4 * H + H, 4 * H, w.Hin + (size_t)b1 * H, H, G[P_CE_WHH], H);
    linear_wgrad(st, n_op, G3, XP, w.gc, 4 * H, w.XnS, XP, w.dWihP[0], XP);
  }
  unpad_add_wih(st, w.dWihP[0], SX, G[P_CE_WIH]);
  unpad_add_wih(st, w.dWihP[1], SX, G[P_LE_WIH]);
  unpad_add_wih(st, w.dWihP[2], SX0, G[P_RE_WIH]);
}

int encode_fwd(dx_stream_t st, const float* weights, const Batch& bt, float* mu, float* std_, void* ws,
               size_t ws_bytes, int keep) {
  Arena ar(ws, ws_bytes);
  EncWs w = carve_enc(ar, bt.B, keep != 0, bt.n_levels, bt.level_ptr);
  DX_CHECK(!ar.overflow, "encode_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, ar.off);
  Weights W(weights);
  encode_fwd_impl(st, W, bt, w, mu, std_, keep != 0);
  return check_launch("encode_fwd");
}

}  // namespace dx
