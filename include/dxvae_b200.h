/* dxvae_b200 — C ABI of the B200-native DX-VAE hot path.
 *
 * The reference (HotzingTone/DX-VAE) has no FFI of its own: its hot path is the
 * Python class DXVAE (model.py:10-391) calling stock torch ops on lists of DGL
 * graphs.  This header is the boundary the new build introduces UNDER that class:
 * the Python mirror dxvae_b200/model.py binds these entry points with ctypes and
 * keeps the reference's method surface (encode / reparameterize / decode / loss /
 * forward / train).  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain pointers + explicit sizes; every pointer is a DEVICE pointer unless the
 *     name ends in _host; the caller (torch) owns all buffers, workspace included;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     nothing synchronises unless stated;
 *   - return 0 = ok, non-zero = error; dxvae_last_error() gives the message
 *     (thread-local).  Nothing throws across the ABI.  Process state is limited to per-device set-up flags
 *     (constant tables, kernel attributes), the launch counter, a per-thread table of encoded TMA tensor maps
 *     (a pure function of address / extents / pitch: cached instead of re-encoded per launch) and the thread-local
 *     arithmetic mode / training scope that an entry point sets for its own duration;
 *   - every kernel is launched with programmatic stream serialization and waits (griddepcontrol.wait) before its
 *     first global access, so back-to-back entry points on one stream overlap launch latency, never data
 *     (DX_NO_PDL=1 in the environment launches plainly).  Work enqueued by OTHER code on the same stream keeps the
 *     usual stream order.
 *   - model constants are fixed (7 nodes, X 27, X0 23, H 512, Z 128): kernels are
 *     specialised on them.
 *
 * Layouts (B = graphs in the batch)
 *   weights  flat fp32 blob, dxvae_param_blob_floats() long; tensor k of the 53
 *            state_dict tensors (App. E order, model.py:24-72) lives at
 *            dxvae_param_entry(k).offset, row-major, each offset 64-float aligned.
 *   Xg       (B,7,27) fp32   graph-major node features  = stack of g.ndata['X']
 *   Pg       (B,7,21) fp32   graph-major raw parameters = stack of g.ndata['params']
 *   Xn       (7,B,32) fp32   node-major, zero padded to 32 columns (what kernels read)
 *   cls      (14,B)  int32   class labels: row0 lfw, row1 alg, rows 2..7 lc(op1..6),
 *                            rows 8..13 rc(op1..6)            (model.py:307-308,327-328)
 *   adj      (B)     uint64  bit (src*7+dst) set iff edge src->dst
 */
#ifndef DXVAE_B200_H
#define DXVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DXVAE_ABI_VERSION 4
#define DXVAE_N_NODES 7
#define DXVAE_N_PARAMS 21
#define DXVAE_SIZE_X 27
#define DXVAE_SIZE_X0 23
#define DXVAE_SIZE_H 512
#define DXVAE_SIZE_Z 128
#define DXVAE_XPAD 32
#define DXVAE_N_TENSORS 53

int dxvae_abi_version(void);
const char* dxvae_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long dxvae_launch_count(void);
/* per-launch CUDA-event timing of the GEMM kernel family (bench.py roofline): totals per tile
 * class (0: FP32 128x128 tiles, 1: FP32 64x64, 2: tcgen05 TF32 / 3xTF32) of device ms, executed flops
 * (2MNK) and launches; each output array has 3 entries. */
void dxvae_prof_begin(int max_launches);
void dxvae_prof_end(double* ms3, double* flops3, long long* n3);

/* ---- parameter blob (state_dict of model.py:24-72) ------------- */
typedef struct {
  const char* name;  /* state_dict key, e.g. "combin_encode.weight_ih" */
  int64_t offset;    /* in floats, into the flat blob */
  int32_t rows;      /* 2-D: rows; 1-D: length */
  int32_t cols;      /* 2-D: cols; 1-D: 0 */
} dxvae_param_entry_t;

int64_t dxvae_param_blob_floats(void);                 /* padded length of the blob */
int64_t dxvae_param_count(void);                       /* 12,083,541 real parameters */
int dxvae_param_entry(int k, dxvae_param_entry_t* out); /* k in [0,53) */

/* ---- batcher (replaces the per-graph DGL queries of model.py:164-177,189-191,
 *      272-280 and the graph construction of dxdata.py:174-312) ------------------- */

/* Host batcher: COO edge lists -> adjacency masks, CSR by destination over flat node
 * ids b*7+v (sources ascending), per-edge feedback marks (0 forward src>dst, 1 back
 * edge src<dst, 2 self-loop), encode level per node and the level schedule (operator
 * rows v*B+b grouped by level, ascending).  All pointers are HOST pointers.
 * edge_ptr has B+1 entries; indices/eflags need edge_ptr[B] entries; level_ptr needs 16 (8 offsets, then at
 * [8 + L] the number of leading rows of level L that a feedback back-edge arrives at);
 * level_rows needs 6*B.  *n_levels receives the number of operator levels. */
int dxvae_batch_build_host(int64_t B, const int32_t* edge_ptr_host, const int8_t* src_host, const int8_t* dst_host,
                           uint64_t* adj_host, int32_t* indptr_host, int32_t* indices_host, uint8_t* eflags_host,
                           uint8_t* level_host, int32_t* level_ptr_host, int32_t* level_rows_host,
                           int32_t* n_levels);

/* Device level schedule from adjacency masks (same outputs as the host batcher's
 * level / level_ptr / level_rows).  level_ptr (16 ints, same layout) is written on the device AND
 * copied to level_ptr_host (pinned or pageable) — this call synchronises the stream. */
int dxvae_batch_schedule(int64_t B, const uint64_t* adj, uint8_t* level, int32_t* level_ptr, int32_t* level_rows,
                         int32_t* level_ptr_host, void* workspace, size_t workspace_bytes, void* stream);

/* Decoder step schedule for teacher forcing (model.py:311-358): the loss replays 21 (vi,vj)
 * re-propagates per graph, but a re-propagate only changes node vi when that step adds an edge
 * (vj->vi or vi->vj); otherwise it recomputes the same state.  step t = vi*(vi-1)/2 + (vi-1-vj).
 * step_rows[step_ptr[t]..step_ptr[t+1]) = ascending ids of the graphs active at step t; passing
 * the schedule to dxvae_elbo_step / dxvae_loss_step makes the identity steps free (results are
 * the same function of the inputs; NULL runs every step on every graph as the reference does).
 * Lists 21..26 hold the graphs with a self-loop on node vi = 1..6: the second propagate of a new node
 * (model.py:337, x_loop = selfloop * x) differs from the first only on those graphs.
 * Lists 27..32 hold the graphs in which node x = 0..5 has an edge to a higher node (a feedback back-edge leaves x):
 * only those rows need the "in" half of x's gate / mapper projections in the decoder.
 * step_ptr: 34 ints; step_rows: up to 33*B ints.  The device form synchronises the stream to
 * return step_ptr_host. */
int dxvae_batch_steps(int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows, int32_t* step_ptr_host,
                      void* workspace, size_t workspace_bytes, void* stream);
int dxvae_batch_steps_host(int64_t B, const uint64_t* adj_host, int32_t* step_ptr_host, int32_t* step_rows_host);

/* Graph-major reference tensors -> what the kernels read. */
int dxvae_pack_graphs(int64_t B, const float* Xg, const float* Pg, float* Xn, int32_t* cls, void* stream);
/* The same for the B rows idx[b] (device int64 list; NULL = identity) of a larger graph-major store Xg / Pg / adj_g — how
 * a training loop draws a shuffled batch from its dataset (model.py:380-382).  Xg, Pg and adj_g may point to PINNED HOST
 * memory (cudaHostAlloc / torch pin_memory: mapped into the device's address space): the gather then is the host-to-device
 * transfer of the batch.  cls and adj may be NULL. */
int dxvae_pack_graphs_indexed(int64_t B, const int64_t* idx, const float* Xg, const float* Pg, const uint64_t* adj_g,
                              float* Xn, int32_t* cls, uint64_t* adj, void* stream);
/* Node-major decode outputs -> graph-major (B,7,27) / (B,7,21). */
int dxvae_unpack_graphs(int64_t B, const float* Xn, const float* Pn, float* Xg, float* Pg, void* stream);

/* dxdata.py:174-312 (_make_graph) on the device: packed 128-byte DX7 voices ->
 * Xn, cls, adj (from the 32-entry DX_ALGO table, dxdata.py:140-171) and, when Pg is
 * non-NULL, the graph-major params (B,7,21). */
int dxvae_voices_to_graphs(int64_t B, const uint8_t* voices, float* Xn, int32_t* cls, uint64_t* adj, float* Xg,
                           float* Pg, void* stream);
/* dxdata.py:341-397 (graph_to_syx): params (B,7,21) fp32 -> B*128 packed voice bytes
 * (header/name/trailer bytes are added by the host wrapper). */
int dxvae_pack_syx(int64_t B, const float* Pg, uint8_t* voices, void* stream);

/* ---- arithmetic of the dense products ---------------------------------------------- *
 * DXVAE_PREC_FP32: FP32 FFMA kernels everywhere (reference-tolerance parity; the only mode of
 *                  greedy decode, whose discrete outputs must match the reference exactly).
 * DXVAE_PREC_TF32: eligible products (rows >= 128, N >= 64, K >= 32, no row gather) run on the
 *                  tcgen05 tensor cores with TF32 inputs / FP32 accumulation; looser, stated
 *                  tolerance.
 * DXVAE_PREC_3XTF32: FP32-accurate products on the tensor cores, every entry point and every operand form (forward,
 *                  dgrad, wgrad).  The kernel itself splits each landed operand tile into tf32 hi + lo parts in
 *                  shared memory and issues hi*hi + hi*lo + lo*hi; the FP32 partial sums are drained from TMEM into
 *                  registers every 32 k (the tensor core accumulates with truncation, which would otherwise bias a
 *                  long reduction).  Measured product error 3e-7..6e-7 of max|C| (the FFMA kernels: ~1e-6);
 *                  a training step meets the reference tolerances of DXVAE_PREC_FP32. */
enum { DXVAE_PREC_FP32 = 0, DXVAE_PREC_TF32 = 1, DXVAE_PREC_3XTF32 = 2 };

/* ---- workspace sizes ------------------------------------------------------------ */
enum { DXVAE_OP_ENCODE = 0, DXVAE_OP_DECODE = 1, DXVAE_OP_TRAIN = 2, DXVAE_OP_SCHEDULE = 3,
       DXVAE_OP_ENCODE_TRAIN = 4, DXVAE_OP_LOSS = 5 };
size_t dxvae_workspace_bytes(int op, int64_t B);
/* The same for a batch whose schedules are known (all host arrays, either may be NULL = worst case):
 *   level_ptr_host (n_levels + 1, dxvae_batch_build_host / dxvae_batch_schedule): the encoder's per-level temporaries
 *     are sized for the largest level instead of 6B rows;
 *   step_ptr_host (dxvae_batch_steps*, DXVAE_OP_TRAIN / DXVAE_OP_LOSS): the per-step activations of the teacher-forced
 *     decoder are kept for the ACTIVE graphs of each step only.
 * Never more than dxvae_workspace_bytes; about 0.65x for training on Dexed-like topologies.  Entry points called with
 * those schedules accept either size. */
size_t dxvae_workspace_bytes_sched(int op, int64_t B, int32_t n_levels, const int32_t* level_ptr_host,
                                   const int32_t* step_ptr_host);

/* ---- encode (model.py:200-212; _propagate :151-198 with encode=True) ------------- *
 * level_ptr_host (n_levels+1 ints, HOST) and level_rows (DEVICE) come from the batcher.
 * level_rare_host (HOST, n_levels ints, or NULL): the batcher orders every level with the rows a feedback
 * back-edge ARRIVES at first and reports how many there are (level_ptr[8 + L]); only those rows need the "out"
 * half of their gate / mapper projections, so the encoder computes that half on the prefix alone.  NULL (or a
 * schedule that is not ordered that way) computes both halves for every row; results are identical.
 * Outputs mu, std (B,128).  With keep=1 the workspace retains what encode_bwd needs
 * (workspace must then be the DXVAE_OP_TRAIN one). */
int dxvae_encode_fwd(const float* weights, int64_t B, const float* Xn, const uint64_t* adj, int32_t n_levels,
                     const int32_t* level_ptr_host, const int32_t* level_rows, const int32_t* level_rare_host, float* mu,
                     float* std_, void* workspace, size_t workspace_bytes, int keep, int precision, void* stream);

/* ---- reparameterise (model.py:284, Normal.rsample): z = mu + std*eps -------------- */
int dxvae_reparameterize(int64_t n, const float* mu, const float* std_, const float* eps, float* z, void* stream);

/* ---- greedy decode (model.py:214-253, quantisers :87-149) ------------------------ *
 * z (B,128) -> Xg (B,7,27), Pg (B,7,21), adj (B).  margins (optional, may be NULL): (B,2) fp32,
 * [b][0] = min |logit| over the 48 edge decisions of graph b (sigmoid > 0.5 flips at logit 0),
 * [b][1] = min distance, in logit units, of any parameter logit to the point where its quantiser would decide
 * otherwise (rounding tie of _q_lin / _q_log, sigmoid at 0.5 of _q_bool, arg-max gap of _q_prob; model.py:87-107),
 * for tie-aware parity checks.  After each pair of edge decisions only the graphs that gained an edge
 * re-propagate; their count is read back to size the next launches, so this call
 * synchronises the stream (21 small copies per call) and must not be stream-captured. */
int dxvae_decode_greedy(const float* weights, int64_t B, const float* z, float* Xg, float* Pg, uint64_t* adj,
                        float* margins, void* workspace, size_t workspace_bytes, int precision, void* stream);

/* ---- teacher-forced ELBO (model.py:270-367) + backward (model.py:385) ------------ *
 * One call = encode_fwd + loss_fwd (+ backward of both when grads != NULL).
 * eps (B,128) is the injected N(0,1) noise.  loss5 (5 floats, device) receives
 * (total, loss_X0, loss_Xi, loss_E, kld*w_kld) of model.py:367 over the B graphs,
 * each scaled by inv_batch * B.  `inv_batch` is 1/(global batch): every term is a
 * batch mean (model.py:303-365), so data-parallel ranks pass 1/(B*world) and sum
 * loss5 / grads across ranks.  grads: flat blob, same layout as weights, ACCUMULATED
 * into (caller zeroes it).  decoder_done_event (cudaEvent_t or NULL) is recorded on the stream once the decoder's
 * backward has been issued: the gradient range of the decoder-only tensors (combin_decode.weight_ih up to, not
 * including, gate.0.weight) is final from then on, so a data-parallel caller can reduce it while the encoder's
 * backward still runs. */
int dxvae_elbo_step(const float* weights, int64_t B, const float* Xn, const int32_t* cls, const uint64_t* adj,
                    int32_t n_levels, const int32_t* level_ptr_host, const int32_t* level_rows,
                    const int32_t* level_rare_host, const float* eps,
                    float w_env, float w_frq, float w_kld, float inv_batch, float* loss5, float* mu_out,
                    float* std_out, float* grads, void* workspace, size_t workspace_bytes, int precision,
                    const int32_t* step_ptr_host, const int32_t* step_rows, void* decoder_done_event, void* stream);

/* Split form of dxvae_elbo_step, for DXVAE.encode(G) followed by DXVAE.loss(q, G)
 * (model.py:370-371).  encode_fwd(keep=1, workspace of DXVAE_OP_ENCODE_TRAIN) leaves the
 * encoder's activations in its workspace; loss_step (workspace DXVAE_OP_LOSS) consumes
 * mu/std and returns the loss terms, decoder gradients (accumulated into grads) and
 * dL/dmu, dL/dstd; encode_bwd then finishes the chain from the SAME encoder workspace. */
int dxvae_loss_step(const float* weights, int64_t B, const float* Xn, const int32_t* cls, const uint64_t* adj,
                    const float* mu, const float* std_, const float* eps, float w_env, float w_frq, float w_kld,
                    float inv_batch, float* loss5, float* grads, float* dmu, float* dstd, void* workspace,
                    size_t workspace_bytes, int precision, const int32_t* step_ptr_host, const int32_t* step_rows,
                    void* stream);
int dxvae_encode_bwd(const float* weights, int64_t B, const float* Xn, const uint64_t* adj, int32_t n_levels,
                     const int32_t* level_ptr_host, const int32_t* level_rows, const int32_t* level_rare_host,
                     const float* std_, const float* dmu,
                     const float* dstd, float* grads, void* workspace, size_t workspace_bytes, int precision,
                     void* stream);

/* ---- optimiser (model.py:375,386: torch.optim.AdamW defaults) -------------------- */
int dxvae_adamw_step(int64_t n, float* weights, const float* grads, float* exp_avg, float* exp_avg_sq, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                     void* stream);

/* ---- low-level pieces exported for unit tests ------------------------------------ *
 * variant 0: C[M,N] = act(A[M,K] W[N,K]^T + bias) (act: 0 none, 1 relu, 2 tanh, 4 softplus);
 * 1: dgrad C[M,K] (+)= A[M,N] W[N,K]; 2: wgrad C[N,K] += A[M,N]^T B[M,K]; +16: TF32 tensor-core path; +32: 3xTF32;
 * 64: forward with bf16 operands (A, Bm point at bf16 data) */
int dxvae_test_gemm(int variant, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* Bm,
                    int64_t ldb, float* C, int64_t ldc, const float* bias, int act, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DXVAE_B200_H */
