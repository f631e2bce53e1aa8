// C ABI (include/dxvae_b200.h) and the composite entry points.
#include <stdarg.h>

#include "../../include/dxvae_b200.h"
#include "dx_engine.h"

namespace dx {

std::atomic<long long> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct TrainWs { EncWs e; DecWs d; float *mu, *sd, *dmu, *dsd; };

static TrainWs carve_train(Arena& ar, int64_t B, int n_levels, const int32_t* level_ptr, const int32_t* step_ptr) {
  TrainWs t;
  t.e = carve_enc(ar, B, true, n_levels, level_ptr);
  t.d = carve_dec(ar, B, true, step_ptr);
  t.mu = ar.take<float>((size_t)B * Z); t.sd = ar.take<float>((size_t)B * Z);
  t.dmu = ar.take<float>((size_t)B * Z); t.dsd = ar.take<float>((size_t)B * Z);
  return t;
}

size_t workspace_bytes(int op, int64_t B, int n_levels, const int32_t* level_ptr, const int32_t* step_ptr) {
  Arena ar(nullptr, (size_t)-1);
  switch (op) {
    case DXVAE_OP_ENCODE: carve_enc(ar, B, false, n_levels, level_ptr); break;
    case DXVAE_OP_DECODE: carve_dec(ar, B, false); break;
    case DXVAE_OP_TRAIN: carve_train(ar, B, n_levels, level_ptr, step_ptr); break;
    case DXVAE_OP_SCHEDULE: ar.take<int32_t>((size_t)72 * ((B + 1023) / 1024)); break;
    case DXVAE_OP_ENCODE_TRAIN: carve_enc(ar, B, true, n_levels, level_ptr); break;
    case DXVAE_OP_LOSS: carve_dec(ar, B, true, step_ptr); break;
    default: return 0;
  }
  size_t need = ar.off;
  if ((op == DXVAE_OP_TRAIN || op == DXVAE_OP_LOSS) && step_ptr == nullptr && B > 0 && B <= small_batch_max()) {
    // Worst case over schedules: the small-batch schedule of the decoder (dx_decoder.cu, p1_batched) keeps extra per-node /
    // per-step buffers, and only when it runs on a compacted schedule — so the bound is the larger of the dense replay
    // (above) and a compacted schedule with every graph in every list (carve_dec grows with every list's row count).
    int32_t dense[NLIST + 1];
    for (int i = 0; i <= NLIST; ++i) dense[i] = (int32_t)(i * B);
    Arena ad(nullptr, (size_t)-1);
    if (op == DXVAE_OP_TRAIN) carve_train(ad, B, n_levels, level_ptr, dense); else carve_dec(ad, B, true, dense);
    if (ad.off > need) need = ad.off;
  }
  return need + 256;
}

// model.py:369-372 forward (= encode + loss) and model.py:385 backward, one call.
int elbo_step(dx_stream_t st, const float* weights, const Batch& bt, const float* eps, LossW lw, float* loss5,
              float* mu_out, float* std_out, float* grads, void* ws, size_t ws_bytes, int precision, void* dec_done_event) {
  const int B = (int)bt.B;
  PrecisionScope prec(precision);
  FwdSplitScope fsplit(true, few_rows_for_batch(bt.B));   // training: few-row forward products may split their reduction (dx_gemm.h)
  Arena ar(ws, ws_bytes);
  TrainWs t = carve_train(ar, bt.B, bt.n_levels, bt.level_ptr, bt.step_ptr);
  DX_CHECK(!ar.overflow, "elbo_step: workspace too small (%zu < %zu bytes)", ws_bytes, ar.off);
  Weights W(weights);
  encode_fwd_impl(st, W, bt, t.e, t.mu, t.sd, true);
  reparameterize(st, (int64_t)B * Z, t.mu, t.sd, eps, t.d.z);
  zero_async(st, t.d.rowloss, sizeof(float) * 4 * (size_t)B);
  DecIO io{true, &bt, lw, nullptr, nullptr};
  if (grads) { Weights Gf(grads); io.dW2 = Gf[P_E_W2]; io.db2 = Gf[P_E_B2]; io.db0 = Gf[P_E_B0]; }
  decode_fwd_impl(st, W, B, t.d.z, t.d, io);
  kld_rows(st, B, t.mu, t.sd, lw, t.d.rowloss);
  loss_reduce(st, B, t.d.rowloss, loss5);
  if (mu_out) copy_async(st, mu_out, t.mu, sizeof(float) * (size_t)B * Z);
  if (std_out) copy_async(st, std_out, t.sd, sizeof(float) * (size_t)B * Z);
  if (grads) {
    Weights G(grads);
    decode_bwd_impl(st, W, G, B, t.d.z, t.d, bt, lw);
#ifndef DX_EMU
    // the decoder-only tensors (combin_decode .. h_to_edge) receive no further gradient: a data-parallel caller may
    // start reducing that range while the encoder backward runs
    if (dec_done_event) cudaEventRecord((cudaEvent_t)dec_done_event, st);
#endif
    latent_bwd(st, B, t.mu, t.sd, eps, t.d.dz, lw, t.dmu, t.dsd);
    encode_bwd_impl(st, W, G, bt, t.e, t.dmu, t.dsd, t.sd);
  }
  return check_launch("elbo_step");
}

// Split form of elbo_step for callers that hold q(z|G) between the two halves
// (DXVAE.encode(...) followed by DXVAE.loss(q, ...), model.py:370-371).
int loss_step(dx_stream_t st, const float* weights, const Batch& bt, const float* mu, const float* sd, const float* eps,
              LossW lw, float* loss5, float* grads, float* dmu, float* dsd, void* ws, size_t ws_bytes, int precision) {
  const int B = (int)bt.B;
  PrecisionScope prec(precision);
  FwdSplitScope fsplit(true, few_rows_for_batch(bt.B));   // training: few-row forward products may split their reduction (dx_gemm.h)
  Arena ar(ws, ws_bytes);
  DecWs d = carve_dec(ar, bt.B, true, bt.step_ptr);
  DX_CHECK(!ar.overflow, "loss_step: workspace too small (%zu < %zu bytes)", ws_bytes, ar.off);
  Weights W(weights);
  reparameterize(st, (int64_t)B * Z, mu, sd, eps, d.z);
  zero_async(st, d.rowloss, sizeof(float) * 4 * (size_t)B);
  DecIO io{true, &bt, lw, nullptr, nullptr};
  if (grads) { Weights Gf(grads); io.dW2 = Gf[P_E_W2]; io.db2 = Gf[P_E_B2]; io.db0 = Gf[P_E_B0]; }
  decode_fwd_impl(st, W, B, d.z, d, io);
  kld_rows(st, B, mu, sd, lw, d.rowloss);
  loss_reduce(st, B, d.rowloss, loss5);
  if (grads) {
    DX_CHECK(dmu && dsd, "loss_step: dmu/dstd required with grads");
    Weights G(grads);
    decode_bwd_impl(st, W, G, B, d.z, d, bt, lw);
    latent_bwd(st, B, mu, sd, eps, d.dz, lw, dmu, dsd);
  }
  return check_launch("loss_step");
}

int encode_bwd(dx_stream_t st, const float* weights, const Batch& bt, const float* sd, const float* dmu,
               const float* dsd, float* grads, void* ws, size_t ws_bytes, int precision) {
  PrecisionScope prec(precision);
  FwdSplitScope fsplit(true, few_rows_for_batch(bt.B));   // training: few-row forward products may split their reduction (dx_gemm.h)
  Arena ar(ws, ws_bytes);
  EncWs e = carve_enc(ar, bt.B, true, bt.n_levels, bt.level_ptr);
  DX_CHECK(!ar.overflow, "encode_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, ar.off);
  Weights W(weights), G(grads);
  encode_bwd_impl(st, W, G, bt, e, dmu, dsd, sd);
  return check_launch("encode_bwd");
}

int decode_greedy(dx_stream_t st, const float* weights, int64_t B64, const float* z, float* Xg, float* Pg,
                  uint64_t* adj, float* margins, void* ws, size_t ws_bytes, int precision) {
  const int B = (int)B64;
  PrecisionScope prec(precision);
  Arena ar(ws, ws_bytes);
  DecWs w = carve_dec(ar, B64, false);
  DX_CHECK(!ar.overflow, "decode_greedy: workspace too small (%zu < %zu bytes)", ws_bytes, ar.off);
  Weights W(weights);
  zero_async(st, adj, sizeof(uint64_t) * (size_t)B);
  if (margins) foreach (st, (int64_t)2 * B, [=] DX_HD(int64_t i) { margins[i] = 3.0e38f; });
  DecIO io{false, nullptr, LossW{0, 0, 0, 0}, adj, margins};
  decode_fwd_impl(st, W, B, z, w, io);
  unpack_graphs(st, B, w.Xd, w.Pn, Xg, Pg);
  return check_launch("decode_greedy");
}

}  // namespace dx

// =============================================================================================
using namespace dx;
#define DX_ST(s) ((dx_stream_t)(s))
#define DX_BATCH_OK(B) DX_CHECK((B) > 0 && (int64_t)7 * (B) < (1ll << 31), "batch size %lld out of range", (long long)(B))

extern "C" {

int dxvae_abi_version(void) { return DXVAE_ABI_VERSION; }
const char* dxvae_last_error(void) { return g_err; }
long long dxvae_launch_count(void) { return g_launches.load(); }
void dxvae_prof_begin(int max_launches) { prof_begin(max_launches); }
void dxvae_prof_end(double* ms, double* flops, long long* n) { prof_end(ms, flops, n); }

int64_t dxvae_param_blob_floats(void) { return param_blob_floats(); }
int64_t dxvae_param_count(void) {
  int64_t n = 0;
  for (int k = 0; k < P_COUNT; ++k) n += param_numel(k);
  return n;
}
int dxvae_param_entry(int k, dxvae_param_entry_t* out) {
  DX_CHECK(k >= 0 && k < P_COUNT && out, "param_entry: index %d out of range", k);
  out->name = kParams[k].name; out->offset = offsets().o[k]; out->rows = kParams[k].rows; out->cols = kParams[k].cols;
  return 0;
}

int dxvae_batch_build_host(int64_t B, const int32_t* edge_ptr_host, const int8_t* src_host, const int8_t* dst_host,
                           uint64_t* adj_host, int32_t* indptr_host, int32_t* indices_host, uint8_t* eflags_host,
                           uint8_t* level_host, int32_t* level_ptr_host, int32_t* level_rows_host, int32_t* n_levels) {
  return batch_build_host(B, edge_ptr_host, src_host, dst_host, adj_host, indptr_host, indices_host, eflags_host,
                          level_host, level_ptr_host, level_rows_host, n_levels);
}
int dxvae_batch_schedule(int64_t B, const uint64_t* adj, uint8_t* level, int32_t* level_ptr, int32_t* level_rows,
                         int32_t* level_ptr_host, void* workspace, size_t workspace_bytes, void* stream) {
  return batch_schedule(DX_ST(stream), B, adj, level, level_ptr, level_rows, level_ptr_host, workspace,
                        workspace_bytes);
}
int dxvae_batch_steps(int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows, int32_t* step_ptr_host,
                      void* workspace, size_t workspace_bytes, void* stream) {
  return batch_steps(DX_ST(stream), B, adj, step_ptr, step_rows, step_ptr_host, workspace, workspace_bytes);
}
int dxvae_batch_steps_host(int64_t B, const uint64_t* adj_host, int32_t* step_ptr_host, int32_t* step_rows_host) {
  DX_CHECK(B > 0, "batch_steps_host: empty batch");
  return batch_steps_host(B, adj_host, step_ptr_host, step_rows_host);
}
int dxvae_pack_graphs(int64_t B, const float* Xg, const float* Pg, float* Xn, int32_t* cls, void* stream) {
  DX_CHECK(B > 0, "pack_graphs: empty batch");
  return pack_graphs(DX_ST(stream), B, Xg, Pg, Xn, cls);
}
int dxvae_pack_graphs_indexed(int64_t B, const int64_t* idx, const float* Xg, const float* Pg, const uint64_t* adj_g,
                              float* Xn, int32_t* cls, uint64_t* adj, void* stream) {
  DX_CHECK(B > 0, "pack_graphs_indexed: empty batch");
  return pack_graphs_indexed(DX_ST(stream), B, idx, Xg, Pg, adj_g, Xn, cls, adj);
}
int dxvae_unpack_graphs(int64_t B, const float* Xn, const float* Pn, float* Xg, float* Pg, void* stream) {
  DX_CHECK(B > 0, "unpack_graphs: empty batch");
  return unpack_graphs(DX_ST(stream), B, Xn, Pn, Xg, Pg);
}
int dxvae_voices_to_graphs(int64_t B, const uint8_t* voices, float* Xn, int32_t* cls, uint64_t* adj, float* Xg,
                           float* Pg, void* stream) {
  DX_CHECK(B > 0, "voices_to_graphs: empty batch");
  return voices_to_graphs(DX_ST(stream), B, voices, Xn, cls, adj, Xg, Pg);
}
int dxvae_pack_syx(int64_t B, const float* Pg, uint8_t* voices, void* stream) {
  DX_CHECK(B > 0, "pack_syx: empty batch");
  return pack_syx(DX_ST(stream), B, Pg, voices);
}
size_t dxvae_workspace_bytes(int op, int64_t B) { return workspace_bytes(op, B); }
size_t dxvae_workspace_bytes_sched(int op, int64_t B, int32_t n_levels, const int32_t* level_ptr_host, const int32_t* step_ptr_host) {
  return workspace_bytes(op, B, n_levels, level_ptr_host, step_ptr_host);
}

int dxvae_encode_fwd(const float* weights, int64_t B, const float* Xn, const uint64_t* adj, int32_t n_levels,
                     const int32_t* level_ptr_host, const int32_t* level_rows, const int32_t* level_rare_host, float* mu,
                     float* std_, void* workspace, size_t workspace_bytes, int keep, int precision, void* stream) {
  DX_BATCH_OK(B);
  DX_CHECK(n_levels >= 1 && n_levels <= 6, "encode_fwd: n_levels=%d", n_levels);
  DX_CHECK(precision >= PREC_FP32 && precision <= PREC_3XTF32, "encode_fwd: unknown precision %d", precision);
  Batch bt{B, Xn, nullptr, adj, n_levels, level_ptr_host, level_rows};
  bt.level_rare = level_rare_host;
  PrecisionScope prec(precision);
  return encode_fwd(DX_ST(stream), weights, bt, mu, std_, workspace, workspace_bytes, keep);
}
int dxvae_reparameterize(int64_t n, const float* mu, const float* std_, const float* eps, float* z, void* stream) {
  return reparameterize(DX_ST(stream), n, mu, std_, eps, z);
}
int dxvae_decode_greedy(const float* weights, int64_t B, const float* z, float* Xg, float* Pg, uint64_t* adj,
                        float* margins, void* workspace, size_t workspace_bytes, int precision, void* stream) {
  DX_BATCH_OK(B);
  DX_CHECK(precision == PREC_FP32 || precision == PREC_3XTF32, "decode_greedy: precision must be FP32 or 3XTF32 (discrete outputs)");
  return decode_greedy(DX_ST(stream), weights, B, z, Xg, Pg, adj, margins, workspace, workspace_bytes, precision);
}
int dxvae_elbo_step(const float* weights, int64_t B, const float* Xn, const int32_t* cls, const uint64_t* adj,
                    int32_t n_levels, const int32_t* level_ptr_host, const int32_t* level_rows,
                    const int32_t* level_rare_host, const float* eps,
                    float w_env, float w_frq, float w_kld, float inv_batch, float* loss5, float* mu_out,
                    float* std_out, float* grads, void* workspace, size_t workspace_bytes, int precision,
                    const int32_t* step_ptr_host, const int32_t* step_rows, void* decoder_done_event, void* stream) {
  DX_BATCH_OK(B);
  DX_CHECK(n_levels >= 1 && n_levels <= 6, "elbo_step: n_levels=%d", n_levels);
  DX_CHECK((step_ptr_host == nullptr) == (step_rows == nullptr), "elbo_step: step_ptr_host and step_rows go together");
  DX_CHECK(precision >= PREC_FP32 && precision <= PREC_3XTF32, "elbo_step: unknown precision %d", precision);
  Batch bt{B, Xn, cls, adj, n_levels, level_ptr_host, level_rows};
  bt.level_rare = level_rare_host;
  bt.step_ptr = step_ptr_host; bt.step_rows = step_rows;
  LossW lw{w_env, w_frq, w_kld, inv_batch};
  return elbo_step(DX_ST(stream), weights, bt, eps, lw, loss5, mu_out, std_out, grads, workspace, workspace_bytes,
                   precision, decoder_done_event);
}
int dxvae_loss_step(const float* weights, int64_t B, const float* Xn, const int32_t* cls, const uint64_t* adj,
                    const float* mu, const float* std_, const float* eps, float w_env, float w_frq, float w_kld,
                    float inv_batch, float* loss5, float* grads, float* dmu, float* dstd, void* workspace,
                    size_t workspace_bytes, int precision, const int32_t* step_ptr_host, const int32_t* step_rows,
                    void* stream) {
  DX_BATCH_OK(B);
  DX_CHECK((step_ptr_host == nullptr) == (step_rows == nullptr), "loss_step: step_ptr_host and step_rows go together");
  DX_CHECK(precision >= PREC_FP32 && precision <= PREC_3XTF32, "loss_step: unknown precision %d", precision);
  Batch bt{B, Xn, cls, adj, 0, nullptr, nullptr};
  bt.step_ptr = step_ptr_host; bt.step_rows = step_rows;
  LossW lw{w_env, w_frq, w_kld, inv_batch};
  return loss_step(DX_ST(stream), weights, bt, mu, std_, eps, lw, loss5, grads, dmu, dstd, workspace, workspace_bytes,
                   precision);
}
int dxvae_encode_bwd(const float* weights, int64_t B, const float* Xn, const uint64_t* adj, int32_t n_levels,
                     const int32_t* level_ptr_host, const int32_t* level_rows, const int32_t* level_rare_host,
                     const float* std_, const float* dmu,
                     const float* dstd, float* grads, void* workspace, size_t workspace_bytes, int precision,
                     void* stream) {
  DX_BATCH_OK(B);
  DX_CHECK(n_levels >= 1 && n_levels <= 6, "encode_bwd: n_levels=%d", n_levels);
  DX_CHECK(precision >= PREC_FP32 && precision <= PREC_3XTF32, "encode_bwd: unknown precision %d", precision);
  Batch bt{B, Xn, nullptr, adj, n_levels, level_ptr_host, level_rows};
  bt.level_rare = level_rare_host;
  return encode_bwd(DX_ST(stream), weights, bt, std_, dmu, dstd, grads, workspace, workspace_bytes, precision);
}
int dxvae_adamw_step(int64_t n, float* weights, const float* grads, float* exp_avg, float* exp_avg_sq, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                     void* stream) {
  DX_CHECK(n > 0 && step >= 1, "adamw_step: bad arguments");
  return adamw_step(DX_ST(stream), n, weights, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step,
                    grad_scale);
}
int dxvae_test_gemm(int variant, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* Bm,
                    int64_t ldb, float* C, int64_t ldc, const float* bias, int act, int accumulate, void* stream) {
  // variant 0: y = act(x W^T + b) ; 1: dx (+)= dy W ; 2: dW += dy^T x ; +16: tcgen05 TF32 path ; +32: 3xTF32 ; +128: training scope
  // variant 64: y = act(x W^T + b) with A and Bm pointing at BF16 data (pitches in elements), tcgen05 kind::f16
  if (variant == 64) {
    GemmP g; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.lda = lda; g.ldb = ldb; g.C = C; g.ldc = ldc; g.bias = bias; g.act = act;
    DX_CHECK(tc_gemm_bf16(DX_ST(stream), g, A, Bm), "test_gemm: bf16 product not eligible");
    return check_launch("test_gemm_bf16");
  }
  PrecisionScope prec((variant & 32) ? PREC_3XTF32 : ((variant & 16) ? PREC_TF32 : PREC_FP32));
  FwdSplitScope fsplit((variant & 128) != 0, (variant & 128) ? 1024 : 256);   // +128: as inside the training entry points (few-row forward products may split)
  variant &= 15;
  DX_CHECK(variant >= 0 && variant <= 2, "test_gemm: unknown variant %d", variant);
  if (variant == 0) linear_fwd(DX_ST(stream), (int)M, (int)N, (int)K, A, lda, Bm, ldb, bias, C, ldc, act);
  else if (variant == 1) linear_dgrad(DX_ST(stream), (int)M, (int)N, (int)K, A, lda, Bm, ldb, C, ldc, accumulate);
  else linear_wgrad(DX_ST(stream), (int)M, (int)N, (int)K, A, lda, Bm, ldb, C, ldc);
  return check_launch("test_gemm");
}

}  // extern "C"
