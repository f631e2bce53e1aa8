// FP32 GEMM used by every dense contraction of the hot path (GRU gate products,
// gate/mapper projections, MLP heads and their backward products).
//   C[M,N] (op)= act( sum_{r<K} Aop(i,r) * Bop(r,j) + bias[j] + add[i,j] )
// Operand forms (memory is always row-major with a leading dimension):
//   a_kc : Aop(i,r) = A[arow(i)*lda + r]      (reduction index contiguous)
//   !a_kc: Aop(i,r) = A[arow(r)*lda + i]      (reduction index = memory row)
//   b_kc : Bop(r,j) = B[j*ldb + r]            (the nn.Linear weight form W[n,k])
//   !b_kc: Bop(r,j) = B[brow(r)*ldb + j]
// a_idx / b_idx / c_idx (optional) remap MEMORY ROWS of A / B / C through an index
// list: this is how the level-scheduled encoder runs a step over a row list.
//   forward  y = x W^T     : a_kc=1, b_kc=1
//   dgrad    dx = dy W     : a_kc=1, b_kc=0
//   wgrad    dW += dy^T x  : a_kc=0, b_kc=0, accum=ACC_ATOMIC (split along the batch)
#pragma once
#include <cstdlib>
#include "dx_rt.h"

namespace dx {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_SOFTPLUS = 4,
       ACT_GATE = 8 };   // `add` is not added but gates the result: C = add[i,j] > 0 ? acc : 0  (relu backward fused into dgrad)
enum { ACC_STORE = 0, ACC_ADD = 1, ACC_ATOMIC = 2 };

struct GemmP {
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; int64_t lda = 0; bool a_kc = true; const int* a_idx = nullptr;
  const float* B = nullptr; int64_t ldb = 0; bool b_kc = true; const int* b_idx = nullptr;
  float* C = nullptr; int64_t ldc = 0; const int* c_idx = nullptr;
  const float* bias = nullptr;
  const float* add = nullptr; int64_t ldadd = 0;
  int act = ACT_NONE;
  int accum = ACC_STORE;
};

enum { PREC_FP32 = 0, PREC_TF32 = 1, PREC_3XTF32 = 2 };
// Arithmetic of the dense products issued by this thread: PREC_FP32 = FFMA kernels (bit-stable),
// PREC_TF32 = tcgen05 tensor cores where the shape is eligible, PREC_3XTF32 = tensor cores with error-compensated
// (hi/lo split) operands: FP32-accurate.
void set_precision(int prec);
int get_precision();
struct PrecisionScope {
  int prev;
  explicit PrecisionScope(int p) : prev(get_precision()) { set_precision(p); }
  ~PrecisionScope() { set_precision(prev); }
};
// Training entry points allow FORWARD products of a few rows (M <= 256) to split their reduction over a thread-block
// cluster (dx_tc_gemm.cu, launch_x3k): deterministic, but the summation order differs from the unsplit kernels, and an
// inference result must not depend on how many patches share the batch — so it is scoped, per thread, like the precision.
// The row limit of that path is scoped the same way: 256 by default, 1024 inside a training step of a SMALL batch
// (B <= small_batch_max(): the launch-bound regime, where the batched parameter heads run 6B rows per product).  A large
// batch keeps 256, so which kernel a compacted step of a few hundred rows takes does not change with this switch.
void set_fwd_split(bool on);
bool get_fwd_split();
void set_few_rows(int rows);
int get_few_rows();
inline int small_batch_max() {
  static const int v = [] { const char* e = getenv("DX_HEADS_BATCH_MAX"); return e ? atoi(e) : 4096; }();
  return v;
}
inline int few_rows_for_batch(int64_t B) {
  static const int forced = [] { const char* e = getenv("DX_X3K_ROWS"); return e ? atoi(e) : 0; }();   // (experiments)
  return forced > 0 ? forced : (B <= small_batch_max() ? 1024 : 256);
}
struct FwdSplitScope {
  bool prev; int prev_rows;
  explicit FwdSplitScope(bool on, int few_rows = 256) : prev(get_fwd_split()), prev_rows(get_few_rows()) { set_fwd_split(on); set_few_rows(few_rows); }
  ~FwdSplitScope() { set_fwd_split(prev); set_few_rows(prev_rows); }
};
bool tc_gemm(dx_stream_t s, const GemmP& p, int* tile_n, bool x3 = false);   // dx_tc_gemm.cu; false = not eligible
// PREC_3XTF32: the same kernels with the operand hi/lo split done inside the kernel (shared memory), three MMAs per
// k-step: FP32-accurate products for every operand form (forward, dgrad, wgrad), no operand copies in HBM.

void gemm(dx_stream_t s, const GemmP& p);
void prof_begin(int max_launches);
void prof_end(double* ms, double* flops, long long* n);

// y[M,N] = act(x[M,K] W[N,K]^T + bias)
inline void linear_fwd(dx_stream_t s, int M, int N, int K, const float* x, int64_t ldx, const float* W, int64_t ldw,
                       const float* bias, float* y, int64_t ldy, int act = ACT_NONE, const int* x_idx = nullptr,
                       const int* y_idx = nullptr, const float* add = nullptr, int64_t ldadd = 0) {
  GemmP p; p.M = M; p.N = N; p.K = K; p.A = x; p.lda = ldx; p.a_kc = true; p.a_idx = x_idx;
  p.B = W; p.ldb = ldw; p.b_kc = true; p.C = y; p.ldc = ldy; p.c_idx = y_idx; p.bias = bias; p.add = add;
  p.ldadd = ldadd; p.act = act; gemm(s, p);
}
// dx[M,K] (+)= dy[M,N] W[N,K]
inline void linear_dgrad(dx_stream_t s, int M, int N, int K, const float* dy, int64_t lddy, const float* W,
                         int64_t ldw, float* dx_, int64_t lddx, int accum, const int* dy_idx = nullptr,
                         const int* dx_idx = nullptr, const float* relu_out = nullptr, int64_t ldrelu = 0) {
  GemmP p; p.M = M; p.N = K; p.K = N; p.A = dy; p.lda = lddy; p.a_kc = true; p.a_idx = dy_idx;
  p.B = W; p.ldb = ldw; p.b_kc = false; p.C = dx_; p.ldc = lddx; p.c_idx = dx_idx; p.accum = accum;
  if (relu_out) { p.add = relu_out; p.ldadd = ldrelu; p.act = ACT_GATE; }   // dx = (relu_out > 0) * (dy W)
  gemm(s, p);
}
// dW[N,K] += dy[M,N]^T x[M,K]   (atomic accumulation; caller zero-initialises dW)
inline void linear_wgrad(dx_stream_t s, int M, int N, int K, const float* dy, int64_t lddy, const float* x,
                         int64_t ldx, float* dW, int64_t lddw, const int* dy_idx = nullptr,
                         const int* x_idx = nullptr) {
  GemmP p; p.M = N; p.N = K; p.K = M; p.A = dy; p.lda = lddy; p.a_kc = false; p.a_idx = dy_idx;
  p.B = x; p.ldb = ldx; p.b_kc = false; p.b_idx = x_idx; p.C = dW; p.ldc = lddw; p.accum = ACC_ATOMIC; gemm(s, p);
}
// bf16 K-major operands (groundwork, test entry only): C = act(A16 B16^T + bias), fp32 accumulate and output
bool tc_gemm_bf16(dx_stream_t s, const GemmP& g, const void* A16, const void* B16);
// db[N] += column sums of dy[M,N] (atomic)
void colsum_accum(dx_stream_t s, int M, int N, const float* dy, int64_t lddy, float* db, const int* dy_idx = nullptr);

}  // namespace dx
