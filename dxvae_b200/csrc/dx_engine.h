// Host-side orchestration of the hot path (internal API behind the C ABI).
#pragma once
#include "dx_gemm.h"
#include "dx_kernels.h"

namespace dx {

struct Weights {  // pointers into the flat blob (weights or gradients: same layout)
  float* p[P_COUNT];
  explicit Weights(const float* blob) {
    const Offsets& o = offsets();
    for (int k = 0; k < P_COUNT; ++k) p[k] = const_cast<float*>(blob) + o.o[k];
  }
  float* operator[](int k) const { return p[k]; }
};

struct Batch {       // what the batcher produces (device pointers except level_ptr)
  int64_t B;
  const float* Xn;   // (7,B,32)
  const int32_t* cls;  // (14,B)   (may be null for encode-only)
  const uint64_t* adj;  // (B)
  int n_levels;
  const int32_t* level_ptr;   // HOST, n_levels+1
  const int32_t* level_rows;  // device, 6B
  const int32_t* level_rare = nullptr;  // HOST, n_levels (optional): leading rows of each level that need the "out" projection half
  // optional decoder step schedule (teacher forcing): rows active at each of the 21 (vi,vj) steps, then
  // (lists 21..26) the graphs with a self-loop on node vi = 1..6 (the only rows where P2 differs from P1), then
  // (lists 27..32) the graphs in which node x = 0..5 has an edge to a higher node (rows needing x's "in" projections)
  const int32_t* step_ptr = nullptr;    // HOST, NLIST+1
  const int32_t* step_rows = nullptr;   // device, step_ptr[NLIST] graph ids
};

struct LossW { float w_env, w_frq, w_kld, inv_batch; };


// ---- workspaces (carved from the caller's buffer; see DESIGN.md "HBM layout") ------------
struct EncWs {
  // All per-node buffers are in SCHEDULE ORDER: position p in [0,6B) = index into level_rows (so the
  // rows of one level are contiguous and every GEMM runs on a dense row range), positions
  // [6B,7B) = node 0 of graph b.  pos[v*B+b] maps a node back to its position.
  float *Hin, *Hc, *Hv, *gc, *gl, *Pg, *Pm, *gxc, *gxl, *gh, *XnS;
  int* pos;
  // input weights padded 27/23 -> 32 columns (16-byte rows, so these products and their gradients are
  // TMA-addressable); training only: their gradients, self-loop-masked features for the looper's weight gradient
  float *WihP[3], *dWihP[3], *XnSL;
  float *dH, *dHin, *dPg, *dPm, *dgb, *dgx, *dgxs, *dgh, *dHc, *dsraw;
};
constexpr int LD_L = 64;  // leading dimension of logit buffers (55 / 27 columns used)
constexpr int NSTEP = 21;
constexpr int NLIST = NSTEP + 6 + 6;   // step-schedule lists: 21 edge steps + 6 self-loop row lists + 6 back-edge-source lists
constexpr int LD_E = 4;    // leading dimension of the edge-head logit buffers (1 or 2 columns used; 16-byte rows for TMA)

struct DecWs {
  float *z, *Hinit, *Hd, *Pg, *Pm, *Q, *g_root;
  float *A1[7], *A2[7], *L[7], *dL[7];
  float *gxc[7], *gxl[7], *Hc0[7], *g_c0[7], *g_p1[7], *Hi_p1[7], *g_p2[7], *Hi_p2[7], *ES1[7], *ls[7], *dls[7];
  float *E1[NSTEP], *l2[NSTEP], *dl2[NSTEP], *Hin[NSTEP], *g_c[NSTEP], *Hc[NSTEP], *g_l[NSTEP], *Hi[NSTEP];
  float *gh, *ghl0, *Hrun, *rowloss;
  float *U, *UC, *dHiC;    // compacted teacher forcing: running edge-head product, compact temp, compact dHi
  float *WihP[3], *dWihP[3], *XL, *xc;   // padded input weights / grads (comb, loop, root), masked features (7B,32), compact x rows
  float *xlS, *xiS;     // compacted teacher forcing, backward: the x rows of every active (graph, step) pair, concatenated in step order
                        // (masked by the self-loop flag / plain): the operands of the ONE deferred weight_ih gradient per cell
  // greedy only
  float *Xd, *Pn;
  int *act_rows, *act_cnt; uint8_t* act_flag;   // graphs that gained an edge at the current step (device-compacted)
  // backward temporaries
  float *dHd, *dPg, *dPm, *dQ, *dgb, *dHi, *dHc, *dHin, *dHrun, *dHc0, *dgx, *dgxs, *dgh, *dE1, *dA1, *dA2, *dES1,
      *dHinit, *dz;
  float *dHc06, *dir6, *dES16, *UCS, *dU6, *U6;   // small batches: first propagates of nodes 1..6 as one 6B-row pass (decode_bwd_impl)
};

struct DecIO {
  bool train;
  const Batch* bt;        // train: teacher-forcing inputs
  LossW lw;
  uint64_t* adj_out;      // greedy: adjacency being built
  float* margins;         // greedy, optional: (B,2) = min |edge logit|, min quantiser margin (logit units)
  float* dW2 = nullptr;   // train + compacted steps: gradient slots of h_to_edge.2.{weight,bias}; the fused edge
  float* db2 = nullptr;   //   head accumulates them during the forward pass (NULL: loss only, no gradients)
  float* db0 = nullptr;   // likewise h_to_edge.0.bias (= column sums of every head's pre-activation gradient)
};

// Gate / mapper projections of a node state (gate.0.weight, mapper.0.weight: (H, 2H)), by half:
//   P[:, :H] = h W[:, :H]^T  ("in" half: the neighbour is a predecessor),  P[:, H:] = h W[:, H:]^T  ("out" half).
// Away from feedback back-edges the encoder only ever reads the "in" half and the decoder the "out" half.
enum { HALF_IN = 0, HALF_OUT = 1 };
inline void proj_fwd(dx_stream_t st, int M, const float* h, const float* Wp, float* P, int half) {
  linear_fwd(st, M, H, H, h, H, Wp + half * H, 2 * H, nullptr, P + half * H, 2 * H);
}
inline void proj_dgrad(dx_stream_t st, int M, const float* dP, const float* Wp, float* dh, int half, int accum) {
  linear_dgrad(st, M, H, H, dP + half * H, 2 * H, Wp + half * H, 2 * H, dh, H, accum);
}
inline void proj_wgrad(dx_stream_t st, int M, const float* dP, const float* h, float* dWp, int half) {
  linear_wgrad(st, M, H, H, dP + half * H, 2 * H, h, H, dWp + half * H, 2 * H);
}

EncWs carve_enc(Arena& ar, int64_t B, bool train, int n_levels = 0, const int32_t* level_ptr = nullptr);
DecWs carve_dec(Arena& ar, int64_t B, bool train, const int32_t* step_ptr = nullptr);
void encode_fwd_impl(dx_stream_t st, const Weights& W, const Batch& bt, const EncWs& w, float* mu, float* sd, bool train);
void encode_bwd_impl(dx_stream_t st, const Weights& W, const Weights& G, const Batch& bt, const EncWs& w,
                     const float* dmu, const float* dstd, const float* sd);
void decode_fwd_impl(dx_stream_t st, const Weights& W, int B, const float* z, const DecWs& w, const DecIO& io);
void decode_bwd_impl(dx_stream_t st, const Weights& W, const Weights& G, int B, const float* z, const DecWs& w,
                     const Batch& bt, LossW lw);
void kld_rows(dx_stream_t st, int B, const float* mu, const float* sd, LossW lw, float* rowloss);
void latent_bwd(dx_stream_t st, int B, const float* mu, const float* sd, const float* eps, const float* dz, LossW lw,
                float* dmu, float* dsd);
void loss_reduce(dx_stream_t st, int B, const float* rowloss, float* out);

// input-weight padding helpers (dx_data.cu)
void pad_wih(dx_stream_t st, const float* W, int K, float* Wp);             // (1536,K) -> (1536,32), zero filled
void unpad_add_wih(dx_stream_t st, const float* dWp, int K, float* dW);      // dW[:, :K] += dWp[:, :K]
void mask_features(dx_stream_t st, int64_t rows, int B, const int* row_ids, int row_base, const uint64_t* adj,
                   const float* X, float* XL);                                // XL[m] = selfloop(row) * X[m]

size_t workspace_bytes(int op, int64_t B, int n_levels = 0, const int32_t* level_ptr = nullptr, const int32_t* step_ptr = nullptr);

int encode_fwd(dx_stream_t st, const float* weights, const Batch& bt, float* mu, float* std_, void* ws, size_t ws_bytes,
               int keep);
int elbo_step(dx_stream_t st, const float* weights, const Batch& bt, const float* eps, LossW lw, float* loss5,
              float* mu_out, float* std_out, float* grads, void* ws, size_t ws_bytes, int precision,
              void* dec_done_event = nullptr);
int decode_greedy(dx_stream_t st, const float* weights, int64_t B, const float* z, float* Xg, float* Pg, uint64_t* adj,
                  float* margins, void* ws, size_t ws_bytes, int precision);

// batcher / data-format kernels (dx_data.cu)
int batch_build_host(int64_t B, const int32_t* edge_ptr, const int8_t* src, const int8_t* dst, uint64_t* adj,
                     int32_t* indptr, int32_t* indices, uint8_t* eflags, uint8_t* level, int32_t* level_ptr,
                     int32_t* level_rows, int32_t* n_levels);
int batch_schedule(dx_stream_t st, int64_t B, const uint64_t* adj, uint8_t* level, int32_t* level_ptr,
                   int32_t* level_rows, int32_t* level_ptr_host, void* ws, size_t ws_bytes);
int batch_steps(dx_stream_t st, int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows,
                int32_t* step_ptr_host, void* ws, size_t ws_bytes);
int batch_steps_host(int64_t B, const uint64_t* adj, int32_t* step_ptr, int32_t* step_rows);
int pack_graphs(dx_stream_t st, int64_t B, const float* Xg, const float* Pg, float* Xn, int32_t* cls);
int pack_graphs_indexed(dx_stream_t st, int64_t B, const int64_t* idx, const float* Xg, const float* Pg,
                        const uint64_t* adjg, float* Xn, int32_t* cls, uint64_t* adj);
int unpack_graphs(dx_stream_t st, int64_t B, const float* Xn, const float* Pn, float* Xg, float* Pg);
int voices_to_graphs(dx_stream_t st, int64_t B, const uint8_t* voices, float* Xn, int32_t* cls, uint64_t* adj,
                     float* Xg, float* Pg);
int pack_syx(dx_stream_t st, int64_t B, const float* Pg, uint8_t* voices);
int adamw_step(dx_stream_t st, int64_t n, float* w, const float* g, float* m, float* v, float lr, float b1, float b2,
               float eps, float wd, int64_t step, float gscale);
int reparameterize(dx_stream_t st, int64_t n, const float* mu, const float* sd, const float* eps, float* z);

}  // namespace dx
