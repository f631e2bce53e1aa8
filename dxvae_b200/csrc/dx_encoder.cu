// Level-scheduled encoder: model.py:200-212 (encode) / :151-198 (_propagate, encode=True).
//
// The reference walks v = 6..0 over the whole batch and asks every DGL graph for its
// neighbours from Python.  Here the batcher's level schedule groups operator rows
// (v*B+b) by topological level; each level is ONE fused step over its row list:
//   gated-sum aggregation (feedback back-edges arrive through the "out" half of the
//   projections) -> combiner GRU -> looper GRU (input masked by the self-loop flag)
//   -> gate/mapper projections of the new state, reused by every lower neighbour.
// The root step (node 0 of all graphs) runs last, then the two latent heads.
#include "dx_engine.h"

namespace dx {

EncWs carve_enc(Arena& ar, int64_t B, bool train) {
  EncWs w{};
  const size_t R7 = (size_t)7 * B, R6 = (size_t)6 * B;
  w.Hin = ar.take<float>(R7 * H); w.Hc = ar.take<float>(R7 * H); w.Hv = ar.take<float>(R7 * H);
  w.Pg = ar.take<float>(R7 * 2 * H); w.Pm = ar.take<float>(R7 * 2 * H);
  w.gxc = ar.take<float>(R6 * G3); w.gxl = ar.take<float>(R6 * G3); w.gh = ar.take<float>(R6 * G3);
  if (train) {
    w.gc = ar.take<float>(R7 * 4 * H); w.gl = ar.take<float>(R7 * 4 * H);
    w.dH = ar.take<float>(R7 * H); w.dHin = ar.take<float>(R7 * H);
    w.dPg = ar.take<float>(R6 * 2 * H); w.dPm = ar.take<float>(R6 * 2 * H); w.dgb = ar.take<float>(R6 * H);
    w.dgx = ar.take<float>(R6 * G3); w.dgxs = ar.take<float>(R6 * G3); w.dgh = ar.take<float>(R6 * G3);
    w.dHc = ar.take<float>(R6 * H); w.dsraw = ar.take<float>((size_t)B * Z);
  }
  return w;
}

void encode_fwd_impl(dx_stream_t st, const Weights& W, const Batch& bt, const EncWs& w, float* mu, float* sd,
                     bool train) {
  const int B = (int)bt.B;
  for (int L = 0; L < bt.n_levels; ++L) {
    const int M = bt.level_ptr[L + 1] - bt.level_ptr[L];
    if (M <= 0) continue;
    const int* rows = bt.level_rows + bt.level_ptr[L];
    RowMap rm{M, B, rows, 0};
    if (L > 0) {
      MsgFwd mf{rm, w.Pg, w.Pm, W[P_G_B], bt.adj, w.Hin, 1, -1, 0, 0};
      msg_fwd(st, mf);
    }
    linear_fwd(st, M, G3, SX, bt.Xn, XP, W[P_CE_WIH], SX, nullptr, w.gxc, G3, ACT_NONE, rows);
    linear_fwd(st, M, G3, SX, bt.Xn, XP, W[P_LE_WIH], SX, nullptr, w.gxl, G3, ACT_NONE, rows);
    if (L > 0) linear_fwd(st, M, G3, H, w.Hin, H, W[P_CE_WHH], H, nullptr, w.gh, G3, ACT_NONE, rows);
    CellFwd c1{rm, w.gxc, L > 0 ? w.gh : nullptr, W[P_CE_BIH], W[P_CE_BHH], L > 0 ? w.Hin : nullptr, 1, w.Hc, 1,
               train ? w.gc : nullptr, 1, S_ONE, bt.adj};
    cell_fwd(st, c1);
    linear_fwd(st, M, G3, H, w.Hc, H, W[P_LE_WHH], H, nullptr, w.gh, G3, ACT_NONE, rows);
    CellFwd c2{rm, w.gxl, w.gh, W[P_LE_BIH], W[P_LE_BHH], w.Hc, 1, w.Hv, 1, train ? w.gl : nullptr, 1, S_SELF, bt.adj};
    cell_fwd(st, c2);
    // projections of the finished state: gate.0.weight (512,1024) viewed as (1024,512) rows (2n,2n+1)=(in,out)
    linear_fwd(st, M, 2 * H, H, w.Hv, H, W[P_G_W], H, nullptr, w.Pg, 2 * H, ACT_NONE, rows, rows);
    linear_fwd(st, M, 2 * H, H, w.Hv, H, W[P_M_W], H, nullptr, w.Pm, 2 * H, ACT_NONE, rows, rows);
  }
  // root step: node 0 of every graph (global rows 0..B-1)
  RowMap r0{B, B, nullptr, 0};
  MsgFwd mf{r0, w.Pg, w.Pm, W[P_G_B], bt.adj, w.Hin, 1, -1, 0, 0};
  msg_fwd(st, mf);
  linear_fwd(st, B, G3, SX0, bt.Xn, XP, W[P_RE_WIH], SX0, nullptr, w.gxc, G3);
  linear_fwd(st, B, G3, H, w.Hin, H, W[P_RE_WHH], H, nullptr, w.gh, G3);
  CellFwd cr{r0, w.gxc, w.gh, W[P_RE_BIH], W[P_RE_BHH], w.Hin, 1, w.Hv, 1, train ? w.gc : nullptr, 1, S_ONE, bt.adj};
  cell_fwd(st, cr);
  linear_fwd(st, B, Z, H, w.Hv, H, W[P_MU_W], H, W[P_MU_B], mu, Z);
  linear_fwd(st, B, Z, H, w.Hv, H, W[P_STD_W], H, W[P_STD_B], sd, Z, ACT_SOFTPLUS);
}

// Backward of the above.  dmu, dstd: (B,128).  Accumulates into the gradient blob G.
void encode_bwd_impl(dx_stream_t st, const Weights& W, const Weights& G, const Batch& bt, const EncWs& w,
                     const float* dmu, const float* dstd, const float* sd) {
  const int B = (int)bt.B;
  // softplus': sigmoid(raw) = 1 - exp(-std)
  {
    float* dsraw = w.dsraw;
    foreach (st, (int64_t)B * Z, [=] DX_HD(int64_t i) { dsraw[i] = dstd[i] * (1.f - expf(-sd[i])); });
  }
  linear_dgrad(st, B, Z, H, dmu, Z, W[P_MU_W], H, w.dH, H, ACC_STORE);
  linear_dgrad(st, B, Z, H, w.dsraw, Z, W[P_STD_W], H, w.dH, H, ACC_ADD);
  linear_wgrad(st, B, Z, H, dmu, Z, w.Hv, H, G[P_MU_W], H);
  linear_wgrad(st, B, Z, H, w.dsraw, Z, w.Hv, H, G[P_STD_W], H);
  colsum_accum(st, B, Z, dmu, Z, G[P_MU_B]);
  colsum_accum(st, B, Z, w.dsraw, Z, G[P_STD_B]);
  // root cell
  RowMap r0{B, B, nullptr, 0};
  CellBwd cr{r0, w.dH, 1, w.gc, 1, w.Hin, 1, w.dgx, nullptr, w.dgh, w.dHin, S_ONE, bt.adj};
  cell_bwd(st, cr);  // dHin rows 0..B-1 <- dh*z (compact == global for the root rows)
  linear_dgrad(st, B, G3, H, w.dgh, G3, W[P_RE_WHH], H, w.dHin, H, ACC_ADD);
  linear_wgrad(st, B, G3, H, w.dgh, G3, w.Hin, H, G[P_RE_WHH], H);
  linear_wgrad(st, B, G3, SX0, w.dgx, G3, bt.Xn, XP, G[P_RE_WIH], SX0);
  colsum_accum(st, B, G3, w.dgh, G3, G[P_RE_BHH]);
  colsum_accum(st, B, G3, w.dgx, G3, G[P_RE_BIH]);

  for (int L = bt.n_levels - 1; L >= 0; --L) {
    const int M = bt.level_ptr[L + 1] - bt.level_ptr[L];
    if (M <= 0) continue;
    const int* rows = bt.level_rows + bt.level_ptr[L];
    RowMap rm{M, B, rows, 0};
    // gradients reaching the projections of these source rows from every lower neighbour
    MsgBwd mb{rm, w.Pg, w.Pm, W[P_G_B], bt.adj, w.dHin, B, w.dPg, w.dPm, w.dgb, 0, -1, 0, 0};
    msg_bwd(st, mb);
    linear_dgrad(st, M, 2 * H, H, w.dPg, 2 * H, W[P_G_W], H, w.dH, H, ACC_STORE, nullptr, rows);
    linear_dgrad(st, M, 2 * H, H, w.dPm, 2 * H, W[P_M_W], H, w.dH, H, ACC_ADD, nullptr, rows);
    linear_wgrad(st, M, 2 * H, H, w.dPg, 2 * H, w.Hv, H, G[P_G_W], H, nullptr, rows);
    linear_wgrad(st, M, 2 * H, H, w.dPm, 2 * H, w.Hv, H, G[P_M_W], H, nullptr, rows);
    colsum_accum(st, M, H, w.dgb, H, G[P_G_B]);
    // looper
    CellBwd cl{rm, w.dH, 1, w.gl, 1, w.Hc, 1, w.dgx, w.dgxs, w.dgh, w.dHc, S_SELF, bt.adj};
    cell_bwd(st, cl);
    linear_dgrad(st, M, G3, H, w.dgh, G3, W[P_LE_WHH], H, w.dHc, H, ACC_ADD);
    linear_wgrad(st, M, G3, H, w.dgh, G3, w.Hc, H, G[P_LE_WHH], H, nullptr, rows);
    linear_wgrad(st, M, G3, SX, w.dgxs, G3, bt.Xn, XP, G[P_LE_WIH], SX, nullptr, rows);
    colsum_accum(st, M, G3, w.dgh, G3, G[P_LE_BHH]);
    colsum_accum(st, M, G3, w.dgx, G3, G[P_LE_BIH]);
    // combiner
    CellBwd cc{rm, w.dHc, 0, w.gc, 1, L > 0 ? w.Hin : nullptr, 1, w.dgx, nullptr, w.dgh, L > 0 ? w.dHin : nullptr,
               S_ONE, bt.adj};
    cc.dhp_global = 1;
    cell_bwd(st, cc);
    if (L > 0) {
      linear_dgrad(st, M, G3, H, w.dgh, G3, W[P_CE_WHH], H, w.dHin, H, ACC_ADD, nullptr, rows);
      linear_wgrad(st, M, G3, H, w.dgh, G3, w.Hin, H, G[P_CE_WHH], H, nullptr, rows);
    }
    linear_wgrad(st, M, G3, SX, w.dgx, G3, bt.Xn, XP, G[P_CE_WIH], SX, nullptr, rows);
    colsum_accum(st, M, G3, w.dgh, G3, G[P_CE_BHH]);
    colsum_accum(st, M, G3, w.dgx, G3, G[P_CE_BIH]);
  }
}

int encode_fwd(dx_stream_t st, const float* weights, const Batch& bt, float* mu, float* std_, void* ws,
               size_t ws_bytes, int keep) {
  Arena ar(ws, ws_bytes);
  EncWs w = carve_enc(ar, bt.B, keep != 0);
  DX_CHECK(!ar.overflow, "encode_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, ar.off);
  Weights W(weights);
  encode_fwd_impl(st, W, bt, w, mu, std_, keep != 0);
  return check_launch("encode_fwd");
}

}  // namespace dx
