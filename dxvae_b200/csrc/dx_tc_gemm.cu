// tcgen05 / TMA GEMMs for sm_100a: the dense contractions of the path (GRU gate products, gate / mapper and
// edge-head projections, MLP layers and their dgrad / wgrad) on the 5th-generation tensor cores, FP32 accumulation in
// TMEM, operands fp32 in HBM.
//
//   C[M,N] (op)= act( sum_r Aop(i,r) Bop(r,j) + bias[j] + add[i,j] )      (same contract as dx_gemm.h)
//
// Three kernel families (dispatch: tc_gemm() at the end of the file):
//   k_tc_gemm_x3w / k_tc_gemm_x3   PREC_3XTF32 (the default arithmetic): FP32-accurate products — operands split
//                                  hi/lo in shared memory by converter warps, hi*lo + lo*hi + hi*hi per k-step,
//                                  chunked accumulation drained into FP32 registers (see the comment above
//                                  k_tc_gemm_x3).  256x256 pair tiles (cta_group::2), 128- and 32-column tiles.
//   k_tc_gemm2 / k_tc_gemm<BN>     PREC_TF32: plain kind::tf32 (the tensor map's TFLOAT32 type rounds on load); pair
//                                  tiles 256x256 and single-CTA 128 x {256,128,64} tiles, double-buffered TMEM
//                                  accumulator so that the epilogue of one tile overlaps the main loop of the next.
//   k_tc_gemm2<.., BF16>           bf16 K-major operands (unit-tested groundwork, not on the product path).
// Common structure: persistent CTAs (one per SM) walk the output tiles;
//   warp 0   : TMA producer  (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx, L2 prefetch of the next tile)
//   warp 1   : TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
//   epilogue warps: tcgen05.ld 32x32b -> registers -> bias/add/act -> swizzled smem slab -> TMA bulk store / reduce-add,
//              two warps per TMEM lane quadrant, half of the columns each
// Both operand majors are native (no physical transposes): K-major tiles are one 2-D box of [rows x 32 floats];
// MN-major tiles (dgrad's W, wgrad's dy and x) are 32x32 boxes laid out as SWIZZLE_128B_BASE32B MN-major atoms (TMA
// swizzle 128B_ATOM_32B; SBO = 512 B between 4-row k groups, LBO = 4096 B between 32-element MN groups) — the only
// MN-major layout tf32 accepts.  Reduction splits (wgrad, sub-wave dgrads) leave through TMA reduce-add.
//
// Roofline: the tensor pipe (kind::tf32 = 1/2 of the bf16 rate; 1/6 per algorithmic flop in the 3xTF32 mode).
#include "dx_gemm.h"

#ifndef DX_EMU
#include <cuda.h>
#include <stdlib.h>

namespace dx {
namespace {

constexpr int TBM = 128, TBK = 32;           // 32 fp32 = one 128-byte swizzle row
constexpr int A_BYTES = TBM * TBK * 4;       // 16 KB

// 3xTF32 converter warps of the chunked kernels below (k_tc_gemm_x3 / k_tc_gemm_x3w): kind::tf32 reads the top 19 bits of
// each 32-bit operand word, so the raw fp32 tile TMA delivers IS the hi part (hi = v with the low 13 mantissa bits
// dropped).  Converter warps write lo = v - hi (exact in fp32, then rounded to tf32) for every element of a landed stage
// into a second buffer of the same layout — element-wise, so it is valid for K-major and MN-major tiles alike.  No
// operand copies in HBM, no extra L2 traffic: the mode costs 3x the MMA issue and one shared-memory read + write of the
// stage.
constexpr int X3_WARPS = 4;
template <int BN> struct TcCfg {
  static constexpr int B_BYTES = BN * TBK * 4;
  static constexpr int RAW = A_BYTES + B_BYTES;                    // bytes TMA delivers per stage
  static constexpr int STAGE = RAW;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int THREADS = 320;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 8 * 4096 /*store staging*/ + 256 /*barriers*/;
};

struct TcParams {
  int M, N, K;
  float* C; int64_t ldc; const int* c_idx;
  const float* bias; const float* add; int64_t ldadd;
  int act, accum, k_chunk;
  int tma_store;    // 1: epilogue stores through TMA (bulk tensor store / reduce-add)
  int add_tma;      // 1: the `add` matrix tile is prefetched into the staging slab by TMA
  int n_fast;       // tile order: 1 = the N tiles of one M panel are adjacent (concurrent CTAs share the big A panel in L2;
                    //             the weight-side operand is small and L2-resident anyway), 0 = M fastest
  int x3_inplace;   // k_tc_gemm_x3: DX_X3_DBG experiment switches (results are wrong when set)
  int x3_chunk;     // (unused)
  int prefetch;     // producers prefetch their next tile's operand boxes into L2 (DX_TC_NO_PREFETCH=1 turns it off)
  long long* dbg;   // optional per-phase clock64() trace of CTA (0,0,0): DX_TC_DEBUG=1
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// L2 prefetch of a tensor-map box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
// SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout)
// layout_type: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (the only layout the
// hardware accepts for MN-major tf32 operands: 4-row x 128-byte atoms, 32-byte swizzle granules).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) /*version*/ | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float tc_act(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_TANH) return tanhf(v);
  if (act == ACT_SOFTPLUS) return softplusf_(v);
  return v;
}

// 3xTF32 converter (warps 10..13, `tid` in [0, 32*X3_WARPS)): lo[i] = raw[i] - tf32_hi(raw[i]) over one landed stage.
// The writes are generic-proxy stores that tcgen05.mma reads through the async proxy: fence before signalling.
template <int NT>
__device__ __forceinline__ void x3_split_stage_t(uint32_t raw, uint32_t lo, int bytes, int tid) {
  // plain (non-volatile) vector accesses, four in flight per thread, so the loads of a batch overlap
  const float4* src = reinterpret_cast<const float4*>(__cvta_shared_to_generic(raw));
  float4* dst = reinterpret_cast<float4*>(__cvta_shared_to_generic(lo));
  const int n = bytes / 16;
  // lo = v - hi is exact in fp32 but has up to 13 significant bits, and the tensor core would TRUNCATE it to 11: round it
  // to the nearest tf32 value here instead (half-up on the magnitude), which halves that residual and removes its bias
  auto lo_of = [](float v) {
    const float l = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    return __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xFFFFE000u);
  };
  int i = tid;
  for (; i + 3 * NT < n; i += 4 * NT) {
    const float4 v0 = src[i], v1 = src[i + NT], v2 = src[i + 2 * NT], v3 = src[i + 3 * NT];
    dst[i] = make_float4(lo_of(v0.x), lo_of(v0.y), lo_of(v0.z), lo_of(v0.w));
    dst[i + NT] = make_float4(lo_of(v1.x), lo_of(v1.y), lo_of(v1.z), lo_of(v1.w));
    dst[i + 2 * NT] = make_float4(lo_of(v2.x), lo_of(v2.y), lo_of(v2.z), lo_of(v2.w));
    dst[i + 3 * NT] = make_float4(lo_of(v3.x), lo_of(v3.y), lo_of(v3.z), lo_of(v3.w));
  }
  for (; i < n; i += NT) { const float4 v = src[i]; dst[i] = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w)); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void x3_split_stage(uint32_t raw, uint32_t lo, int bytes, int tid, bool inplace) {
  (void)inplace;
  x3_split_stage_t<32 * X3_WARPS>(raw, lo, bytes, tid);
}

// One 32-row x 32-column piece of the output (v = this thread's accumulator row, columns gj..gj+31 of the tile row
// block starting at row mq = m0 + 32*q): add / gate / bias / activation, then out through the warp's staging slab.
struct EpiWarp {
  uint32_t slab, abar, nstore; int lane, rsub, cc4; bool vec;
};
__device__ __forceinline__ void epi_cols32(const TcParams& p, const CUtensorMap* tmC, const CUtensorMap* tmAdd, float (&v)[32],
                                           int gj, int mq, EpiWarp& w) {
  const int lane = w.lane, rsub = w.rsub, cc4 = w.cc4;
  const uint32_t slab = w.slab, abar = w.abar;
  uint32_t& nstore = w.nstore;
  const bool vec = w.vec;
  const int gi = mq + lane;                                      // the accumulator row this thread owns
  const bool row_ok = gi < p.M;
  const bool full = gj + 32 <= p.N;                              // warp-uniform: whole chunk in bounds
  const int m0 = mq, q = 0;                                      // (row block base: m0 + q * 32 == mq)
    // the previous bulk store of this warp must have finished reading the slab
    if (nstore > 0) { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); __syncwarp(); }
    if (p.add_tma) {
      // `add` tile (32 rows x 32 cols of this warp) fetched by TMA into the slab it is later stored from
      if (lane == 0) { mbar_expect_tx(abar, 4096); tma_load_2d(tmAdd, slab, abar, gj, m0 + q * 32); }
      mbar_wait(abar, nstore & 1);
      const uint32_t sl = slab + (uint32_t)(lane * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a0, a1, a2, a3;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3)
                     : "r"(sl + (uint32_t)(((j ^ (lane & 7)) << 4))) : "memory");
        if (p.act == ACT_GATE) {
          v[4 * j] = a0 > 0.f ? v[4 * j] : 0.f; v[4 * j + 1] = a1 > 0.f ? v[4 * j + 1] : 0.f;
          v[4 * j + 2] = a2 > 0.f ? v[4 * j + 2] : 0.f; v[4 * j + 3] = a3 > 0.f ? v[4 * j + 3] : 0.f;
        } else { v[4 * j] += a0; v[4 * j + 1] += a1; v[4 * j + 2] += a2; v[4 * j + 3] += a3; }
      }
    } else if (p.add) {
      if (row_ok) {
        const float* ar = p.add + (int64_t)gi * p.ldadd + gj;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(ar + j));
            if (p.act == ACT_GATE) {
              v[j] = a4.x > 0.f ? v[j] : 0.f; v[j + 1] = a4.y > 0.f ? v[j + 1] : 0.f;
              v[j + 2] = a4.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = a4.w > 0.f ? v[j + 3] : 0.f;
            } else { v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gj + j < p.N) { const float a1 = __ldg(ar + j); if (p.act == ACT_GATE) v[j] = a1 > 0.f ? v[j] : 0.f; else v[j] += a1; }
        }
      }
    }
    if (p.bias) {
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gj + j));   // same address in every lane
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (gj + j < p.N) v[j] += __ldg(p.bias + gj + j);
      }
    }
    if (p.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (p.act != ACT_NONE && p.act != ACT_GATE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tc_act(v[j], p.act);
    }
    // Each thread owns one accumulator row; rows go through the 128B-swizzled slab (conflict-free
    // 128-bit accesses both ways).
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t dst = slab + (uint32_t)(lane * 128) + (uint32_t)(((j ^ (lane & 7)) << 4));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                   "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
    }
    if (p.tma_store) {
      // ... and leave as one 4 KB TMA bulk store (plain, or f32 reduce-add for accumulate / split-K
      // modes): bulk stores are not limited by the epilogue warps' outstanding-store budget, unlike
      // STG (measured 12x faster here).
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (p.accum == ACC_STORE) tma_store_2d(tmC, slab, gj, m0 + q * 32);
        else tma_reduce_add_2d(tmC, slab, gj, m0 + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      ++nstore;
    } else {
      // Fallback (row scatter through c_idx, or rows that are not 16-byte aligned): read the slab back
      // row-wise and issue coalesced STG / RED (one warp instruction = 4 rows x 128 contiguous bytes).
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + rsub;
        const int gr = m0 + q * 32 + rr, gc = gj + cc4 * 4;
        float o[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
                     : "r"(slab + (uint32_t)(rr * 128) + (uint32_t)(((cc4 ^ (rr & 7)) << 4))) : "memory");
        if (gr < p.M && gc < p.N) {
          const int64_t crow = p.c_idx ? p.c_idx[gr] : gr;
          float* dst = p.C + crow * p.ldc + gc;
          if (p.accum == ACC_ATOMIC) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gc + e < p.N) atomicAdd(dst + e, o[e]);
          } else if (vec && gc + 3 < p.N) {
            float4 w4 = make_float4(o[0], o[1], o[2], o[3]);
            if (p.accum == ACC_ADD) { const float4 old = *reinterpret_cast<float4*>(dst); w4.x += old.x; w4.y += old.y; w4.z += old.z; w4.w += old.w; }
            *reinterpret_cast<float4*>(dst) = w4;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gc + e < p.N) { if (p.accum == ACC_ADD) dst[e] += o[e]; else dst[e] = o[e]; }
          }
        }
      }
      __syncwarp();
    }
}

// Epilogue of the persistent GEMM kernels (warps 2..9 of a CTA): drains the CTA's 128 accumulator lanes.
template <int BN, class TileFn, class ArriveFn>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, const CUtensorMap* tmC, const CUtensorMap* tmAdd,
                                            uint32_t tmem_base, uint32_t stg_base, uint32_t abar0, uint32_t tfull0,
                                            int t_first, int t_stride, int total, TileFn tile_coords,
                                            ArriveFn arrive_empty, bool trace) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Two warps per TMEM lane quadrant (a warp may only read lanes 32*(warp%4)..+31), each taking
  // half of the tile's columns, so draining an accumulator costs about half a main loop.
  const int ew = warp - 2;                                       // 0..7
  const int q = warp & 3;                                        // TMEM lane quadrant
  const int half = ew >> 2;                                      // which half of the BN columns
  constexpr int HC = BN / 2;
  EpiWarp w{stg_base + (uint32_t)ew * 4096u,                     // one 32x32 fp32 staging slab per warp
            abar0 + 8u * ew,                                     // `add` tile arrival barrier of this warp
            0u, lane, lane >> 3, lane & 7,
            ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.ldc % 4 == 0)};
  uint32_t lt = 0;
  for (int t = t_first; t < total; t += t_stride, ++lt) {
    int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
    const uint32_t as = lt & 1;
    mbar_wait(tfull0 + 8u * as, (lt >> 1) & 1);
    if (trace && threadIdx.x == 64 && lt < 8) p.dbg[208 + 2 * lt] = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = half * HC; c < (half + 1) * HC; c += 32) {
      float v[32];
      tmem_ld32(tacc + (uint32_t)c, v);
      const int gj = n0 + c;
      if (gj >= p.N) continue;                                   // warp-uniform
      epi_cols32(p, tmC, tmAdd, v, gj, m0 + q * 32, w);
    }
    // this warp's TMEM reads of the accumulator are complete: hand it back to the MMA warp
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) arrive_empty(as);
    if (trace && threadIdx.x == 64 && lt < 8) p.dbg[209 + 2 * lt] = clock64();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive the stores
  __syncwarp();
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TcCfg<BN>::THREADS, 1) k_tc_gemm(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB,
                                                    const __grid_constant__ CUtensorMap tmC,
                                                    const __grid_constant__ CUtensorMap tmAdd, const TcParams p) {
  // Persistent: CTA b processes tiles b, b+grid, ... ; two accumulators in TMEM so the epilogue of
  // tile i overlaps the main loop of tile i+1.
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::STAGES;
  constexpr int RAW = Cfg::RAW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // swizzle atoms need 1024-byte alignment
  const uint32_t stg_base = base + S * Cfg::STAGE;                    // 4 warps x 2 x 4 KB store staging
  const uint32_t bars = stg_base + 8 * 4096;                      // full[S], empty[S], tfull[2], tempty[2], slot, add[8]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gm = (p.M + TBM - 1) / TBM, gn = (p.N + BN - 1) / BN;
  const int splits = (p.K + p.k_chunk - 1) / p.k_chunk;
  const int total = gm * gn * splits;
  const bool trace = p.dbg && blockIdx.x == 0;
  if (trace && threadIdx.x == 0) p.dbg[200] = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    for (int w = 0; w < 8; ++w) mbar_init(bars + 8u * (2 * S + 5 + w), 1);   // per-warp `add` tile barriers
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                                      // set-up above overlapped the previous kernel; its results are visible from here on
  if (trace && threadIdx.x == 0) p.dbg[201] = clock64();

  auto tile_coords = [&](int t, int& m0, int& n0, int& kbeg, int& nkb) {
    const int mt = p.n_fast ? (t / gn) % gm : t % gm, nt = p.n_fast ? t % gn : (t / gm) % gn, z = t / (gm * gn);
    m0 = mt * TBM; n0 = nt * BN; kbeg = z * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    nkb = (kend - kbeg + TBK - 1) / TBK;
  };

  if (warp == 0) {
    if (lane == 0) {                                               // ---- TMA producer
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(empty_bar(s), ((it / S) & 1) ^ 1);
          if (trace && it < 32) p.dbg[it] = clock64();
          mbar_expect_tx(full_bar(s), RAW);
          const int k0 = kbeg + kb * TBK;
          const uint32_t sa = base + s * Cfg::STAGE, sb = sa + A_BYTES;
          if (!A_MN) tma_load_2d(&tmA, sa, full_bar(s), k0, m0);
          else
#pragma unroll
            for (int g = 0; g < TBM / 32; ++g) tma_load_2d(&tmA, sa + g * 4096, full_bar(s), m0 + g * 32, k0);
          if (!B_MN) tma_load_2d(&tmB, sb, full_bar(s), k0, n0);
          else
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) tma_load_2d(&tmB, sb + g * 4096, full_bar(s), n0 + g * 32, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                               // ---- MMA issuer
      // instruction descriptor: D=F32, A=B=TF32, majors, N>>3, M>>4   (cute::UMMA::InstrDescriptor)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
        int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
        const uint32_t as = lt & 1;
        mbar_wait(tempty_bar(as), ((lt >> 1) & 1) ^ 1);            // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(full_bar(s), (it / S) & 1);
          if (trace && it < 32) p.dbg[64 + it] = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * Cfg::STAGE, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {                      // UMMA_K = 8 for tf32
            // K-major: 8-row x 128 B atoms, SBO 1024; one UMMA_K = 32 B along the swizzled row.
            // MN-major: 4-row atoms (SBO 512 B), 32-element MN groups 4096 B apart (LBO); UMMA_K = 8 rows.
            const uint32_t oa = A_MN ? sa + k * 1024 : sa + k * 32, ob = B_MN ? sb + k * 1024 : sb + k * 32;
            const uint64_t ad = A_MN ? umma_desc(oa, 4096, 512, 1) : umma_desc(oa, 16, 1024, 2);
            const uint64_t bd = B_MN ? umma_desc(ob, 4096, 512, 1) : umma_desc(ob, 16, 1024, 2);
            umma_tf32(tacc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));                               // frees the smem stage when these MMAs retire
        }
        umma_commit(tfull_bar(as));                                // accumulator complete
      }
    }
  } else if (warp < 10) {                                          // ---- epilogue: warps 2..9
    tc_epilogue<BN>(p, &tmC, &tmAdd, tmem_base, stg_base, bars + 8u * (2 * S + 5), tfull_bar(0), (int)blockIdx.x,
                    (int)gridDim.x, total, tile_coords,
                    [&](uint32_t as) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(as)) : "memory"); },
                    trace);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// =============================================================================================
// 2-CTA variant (cta_group::2) for the large products: a cluster of two CTAs on one TPC computes a
// 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the
// 256 N rows); tcgen05.mma.cta_group::2, issued by the leader CTA only, reads both halves, so the
// L2 -> SM traffic per flop drops by a third against the 128 x 256 single-CTA tile (the single-CTA
// main loop is L2-feed bound: ncu shows the tensor pipe 50 % active).  Both CTAs' TMA loads signal
// the leader's "full" barrier; tcgen05.commit multicasts to both CTAs' "empty" / "accumulator full"
// barriers; each CTA drains its own 128 TMEM lanes and both report to the leader's "accumulator
// empty" barrier.
// =============================================================================================
constexpr int RAW2 = A_BYTES + 128 * TBK * 4;     // 32 KB per CTA per stage from TMA
struct Tc2Cfg {
  static constexpr int S = 6;
  static constexpr int STAGE = RAW2;
  static constexpr int THREADS = 320;
  static constexpr int SMEM = S * STAGE + 1024 + 8 * 4096 + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// arrive on the LEADER CTA's copy of a barrier (remote arrive from the peer, local from the leader itself)
// Default semantics (.release.cta), as cutlass::arch::ClusterBarrier::arrive(cta_id): the explicit `.release.cluster` form
// compiles to MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the arrive, and ncu's source view put 43 % of
// the epilogue warps' samples of k_tc_gemm_x3w on that ERRBAR (profiles/r02a): the "accumulator drained" signal, sent once
// per k-block by 16 warps, was the critical path of the chunk ping-pong.  What the signals order is covered without it:
// TMEM reads by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, the converters' shared-memory writes by
// fence.proxy.async (both issued before the arrive).
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(ra) : "r"(bar));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// TMA load whose completion bytes go to the LEADER CTA's mbarrier (peer bit cleared: cute Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// BF16 (K-major operands only, groundwork for DESIGN.md §7 item 1): the operands are bf16 in HBM; a 128-byte tile row then
// holds 64 k-values and one UMMA (kind::f16) covers K = 16 = the same 32 bytes, so stage bytes, swizzle, descriptors and
// barriers are unchanged — only the k extent of a stage, the instruction kind and the format codes differ.
template <bool A_MN, bool B_MN, bool BF16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Cfg::THREADS, 1)
k_tc_gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAdd, const TcParams p) {
  static_assert(!BF16 || (!A_MN && !B_MN), "bf16 operands: K-major only");
  constexpr int BN = 256, S = Tc2Cfg::S, STAGE2 = Tc2Cfg::STAGE;
  constexpr int TKE = BF16 ? 64 : TBK;                              // k-values per stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = base + S * STAGE2;
  const uint32_t bars = stg_base + 8 * 4096;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int gm2 = (p.M + 255) / 256, gn = (p.N + BN - 1) / BN;
  const int splits = (p.K + p.k_chunk - 1) / p.k_chunk;
  const int total = gm2 * gn * splits;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const bool trace = p.dbg && blockIdx.x == 0;
  if (trace && threadIdx.x == 0) p.dbg[200] = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }   // 8 warps x 2 CTAs
    for (int w = 0; w < 8; ++w) mbar_init(bars + 8u * (2 * S + 5 + w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                              // peer barriers are initialised, TMEM is allocated
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                                      // set-up above overlapped the previous kernel; its results are visible from here on
  if (trace && threadIdx.x == 0) p.dbg[201] = clock64();

  auto tile_coords = [&](int t, int& m0, int& n0, int& kbeg, int& nkb) {
    const int mt = p.n_fast ? (t / gn) % gm2 : t % gm2, nt = p.n_fast ? t % gn : (t / gm2) % gn, z = t / (gm2 * gn);
    m0 = mt * 256 + (int)rank * TBM; n0 = nt * BN; kbeg = z * p.k_chunk;   // this CTA's 128 rows of the pair's tile
    const int kend = min(p.K, kbeg + p.k_chunk);
    nkb = (kend - kbeg + TKE - 1) / TKE;
  };

  if (warp == 0) {
    if (lane == 0) {                                               // ---- TMA producer (both CTAs)
      uint32_t it = 0;
      for (int t = cid; t < total; t += ncl) {
        int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
        const int nb0 = n0 + (int)rank * 128;                      // this CTA's half of the B tile
        // L2 prefetch of this cluster's next tile (see k_tc_gemm_x3w): the clusters sharing a panel ask for it at the
        // same moment, so without it every operand load sees DRAM latency
        int pm0 = 0, pn0 = 0, pk = 0, pnkb = 0;
        const bool pf = p.prefetch && t + ncl < total;
        if (pf) tile_coords(t + ncl, pm0, pn0, pk, pnkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait_cluster(empty_bar(s), ((it / S) & 1) ^ 1);
          if (trace && it < 32) p.dbg[it] = clock64();
          const int k0 = kbeg + kb * TKE;
          if (pf && kb < pnkb) {
            const int q0 = pk + kb * TKE, qn = pn0 + (int)rank * 128;
            if (!A_MN) tma_prefetch_2d(&tmA, q0, pm0);
            else
#pragma unroll
              for (int g = 0; g < 4; ++g) tma_prefetch_2d(&tmA, pm0 + g * 32, q0);
            if (!B_MN) tma_prefetch_2d(&tmB, q0, qn);
            else
#pragma unroll
              for (int g = 0; g < 4; ++g) tma_prefetch_2d(&tmB, qn + g * 32, q0);
          }
          const uint32_t sa = base + s * STAGE2, sb = sa + A_BYTES;
          if (leader) mbar_expect_tx(full_bar(s), 2 * RAW2);       // bytes of both CTAs land on the leader's barrier
          if (!A_MN) tma_load_2d_2sm(&tmA, sa, full_bar(s), k0, m0);
          else
#pragma unroll
            for (int g = 0; g < 4; ++g) tma_load_2d_2sm(&tmA, sa + g * 4096, full_bar(s), m0 + g * 32, k0);
          if (!B_MN) tma_load_2d_2sm(&tmB, sb, full_bar(s), k0, nb0);
          else
#pragma unroll
            for (int g = 0; g < 4; ++g) tma_load_2d_2sm(&tmB, sb + g * 4096, full_bar(s), nb0 + g * 32, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {                                     // ---- MMA issuer (leader CTA only)
      const uint32_t fmt = BF16 ? 1u : 2u;                         // F16F32Format: 1 = BF16, 2 = TF32
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t it = 0, lt = 0;
      for (int t = cid; t < total; t += ncl, ++lt) {
        int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
        const uint32_t as = lt & 1;
        mbar_wait_cluster(tempty_bar(as), ((lt >> 1) & 1) ^ 1);    // both CTAs have drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait_cluster(full_bar(s), (it / S) & 1);
          if (trace && it < 32) p.dbg[64 + it] = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * STAGE2, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {
            const uint32_t oa = A_MN ? sa + k * 1024 : sa + k * 32, ob = B_MN ? sb + k * 1024 : sb + k * 32;
            const uint64_t ad = A_MN ? umma_desc(oa, 4096, 512, 1) : umma_desc(oa, 16, 1024, 2);
            const uint64_t bd = B_MN ? umma_desc(ob, 4096, 512, 1) : umma_desc(ob, 16, 1024, 2);
            if (BF16) umma_f16_2sm(tacc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_tf32_2sm(tacc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(empty_bar(s));                           // frees the stage in both CTAs
        }
        umma_commit_2sm(tfull_bar(as));                            // accumulator complete, both CTAs' epilogues
      }
    }
  } else if (warp < 10) {
    tc_epilogue<BN>(p, &tmC, &tmAdd, tmem_base, stg_base, bars + 8u * (2 * S + 5), tfull_bar(0), cid, ncl, total, tile_coords,
                    [&](uint32_t as) { mbar_arrive_leader(tempty_bar(as)); },   // the LEADER's "accumulator empty" barrier
                    trace);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                              // nobody may still address the peer's smem / TMEM
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// =============================================================================================
// 3xTF32 with chunked accumulation: FP32-accurate products on the tensor cores.
//
// Measured on B200 (tools/x3_bias_probe.py): tcgen05 adds each MMA's partial product into the FP32 TMEM
// accumulator with TRUNCATION toward zero, about one ulp per instruction, so a K = 1024 reduction issued as
// 3 x 128 MMAs into one accumulator comes out 2e-5 short — a bias that is coherent across the 40 dependent
// products of a training step (gradients off by 5e-4).  The error is proportional to the number of MMAs
// accumulated in TMEM at full magnitude, so the kernels keep that number at FOUR:
//   * a "chunk" is ONE k-block (32 k): its 8 cross-term MMAs (hi*lo, lo*hi) are issued FIRST into the fresh chunk
//     accumulator, while it only holds values 2^-11 of the final size (their truncation is then negligible),
//     followed by the 4 hi*hi MMAs;
//   * the epilogue warps drain every finished chunk into FP32 REGISTERS (round-to-nearest adds) while the next
//     chunk runs in the other TMEM buffer.
// k_tc_gemm_x3 (this kernel, 128-column tiles: 128 x 128 per CTA, or 256 x 128 per CTA pair with cta_group::2) and
// k_tc_gemm_x3w (256 x 256 pair tiles, below) issue the SAME instruction sequence per output element, so a
// product's value does not depend on which of them the row count selects: a patch's result is bit-identical
// wherever it sits in a batch (tests/test_gpu_scale.py).
// TMEM: 2 chunk buffers x 128 columns.  Warps: 0 TMA, 1 MMA issue, 2-9 epilogue (64 columns of one lane quadrant
// each: 64 running sums per thread), 10-13 converters (lo = v - tf32(v) over each landed stage, see x3_split_stage).
// =============================================================================================
// BNT = 32: narrow tiles for the thin products of the path (the 27 / 55-column parameter heads, the 32-column padded
// weight_ih gradients): a 128-column tile spends 3/4 of its MMAs on padding there.
template <bool CTA2, int BNT = 128> struct X3Cfg {
  static_assert(BNT == 128 || (BNT == 32 && !CTA2), "tile widths: 128, or 32 on single CTAs");
  static constexpr int BN = BNT;
  static constexpr int BROWS = CTA2 ? BN / 2 : BN;                 // B rows staged by one CTA
  static constexpr int RAW = A_BYTES + BROWS * TBK * 4;            // 24 KB / 32 KB (20 KB for the narrow tile)
  static constexpr int STAGE = 2 * RAW;
  static constexpr int S = (BNT == 32 || CTA2) ? 4 : 3;
  static constexpr int THREADS = 320 + 32 * X3_WARPS;
  static constexpr int SMEM = S * STAGE + 1024 + 8 * 4096 + 256;
};

// KSP (few-row products, M <= 256: see launch_x3k): the reduction is split over the CTAs of a thread-block cluster
// (rank = split index, one output tile per cluster); every CTA runs the unchanged main loop on its k range, parks its
// FP32 partial tile in its own shared memory, and after a cluster barrier each CTA sums a 128/KS-row slice of the tile
// over all ranks through distributed shared memory IN RANK ORDER (deterministic), applies the epilogue and stores it.
template <bool A_MN, bool B_MN, bool CTA2, int BNT = 128, bool KSP = false>
__global__ void __launch_bounds__(X3Cfg<CTA2, BNT>::THREADS, 1)
k_tc_gemm_x3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAdd, const TcParams p) {
  using Cfg = X3Cfg<CTA2, BNT>;
  constexpr int BN = Cfg::BN, S = Cfg::S, RAW = Cfg::RAW, STAGE = Cfg::STAGE, BROWS = Cfg::BROWS;
  constexpr int WC = BN == 128 ? 64 : 32;                          // columns drained by one epilogue warp (narrow tile: warps 2-5 only)
  constexpr int TM = CTA2 ? 256 : 128;                             // rows of the (pair) tile
  constexpr uint32_t NEPI = CTA2 ? 16 : 8;                         // epilogue warps reporting to a leader barrier
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = base + S * STAGE;
  const uint32_t bars = stg_base + 8 * 4096;
  // full[S] empty[S] conv[S] | cfull[2] cempty[2] | tmem slot | add[8]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto conv_bar = [&](int s) { return bars + 8u * (2 * S + s); };
  auto cfull_bar = [&](int b) { return bars + 8u * (3 * S + b); };
  auto cempty_bar = [&](int b) { return bars + 8u * (3 * S + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (3 * S + 4);
  const uint32_t add_bar0 = bars + 8u * (3 * S + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(!KSP || (!CTA2 && BNT == 128), "cluster split-K: single-CTA 128 x 128 tiles");
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int gm = (p.M + TM - 1) / TM, gn = (p.N + BN - 1) / BN;
  const uint32_t ks_rank = KSP ? cluster_ctarank() : 0u, ks_n = KSP ? cluster_nctarank() : 1u;   // split index / splits
  const int splits = KSP ? 1 : (p.K + p.k_chunk - 1) / p.k_chunk;
  const int total = gm * gn * splits;
  const int cid = KSP ? (int)(blockIdx.x / ks_n) : (CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
  const int ncl = KSP ? (int)(gridDim.x / ks_n) : (CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x);   // (KSP: ncl == total, one tile per cluster)
  const int dbg = p.x3_inplace;                                    // DX_X3_DBG experiment switches (results are wrong when set)
  const bool trace = p.dbg && blockIdx.x == 0;                     // DX_TC_DEBUG: clock64() stamps of the first 40 k-blocks of CTA 0
  if (trace && threadIdx.x == 0) p.dbg[250] = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(conv_bar(s), (CTA2 ? 2 : 1) * X3_WARPS); }
    for (int b = 0; b < 2; ++b) { mbar_init(cfull_bar(b), 1); mbar_init(cempty_bar(b), NEPI); }
    for (int w = 0; w < 8; ++w) mbar_init(add_bar0 + 8u * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTA2) cluster_sync_all(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                                      // set-up above overlapped the previous kernel; its results are visible from here on

  auto tile_coords = [&](int t, int& m0, int& n0, int& kbeg, int& nkb) {
    const int mt = p.n_fast ? (t / gn) % gm : t % gm, nt = p.n_fast ? t % gn : (t / gm) % gn;
    const int z = KSP ? (int)ks_rank : t / (gm * gn);
    m0 = mt * TM + (int)rank * TBM; n0 = nt * BN; kbeg = z * p.k_chunk;   // this CTA's 128 rows
    const int kend = min(p.K, kbeg + p.k_chunk);
    nkb = kend > kbeg ? (kend - kbeg + TBK - 1) / TBK : 0;               // (KSP: the last ranks of a short reduction may have none)
  };
  auto wait_leader = [&](uint32_t bar, uint32_t parity) { if (CTA2) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity); };
  auto arrive_leader = [&](uint32_t bar) {
    if (CTA2) mbar_arrive_leader(bar);
    else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  };
  auto commit = [&](uint32_t bar) { if (CTA2) umma_commit_2sm(bar); else umma_commit(bar); };

  if (warp == 0) {
    if (lane == 0) {                                               // ---- TMA producer: bytes land on this CTA's own barrier
      uint32_t it = 0;
      for (int t = cid; t < total; t += ncl) {
        int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
        const int nb0 = n0 + (int)rank * BROWS;                    // this CTA's rows of the B tile
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          wait_leader(empty_bar(s), ((it / S) & 1) ^ 1);
          if (trace && it < 40) p.dbg[it] = clock64();
          mbar_expect_tx(full_bar(s), RAW);
          const int k0 = kbeg + kb * TBK;
          const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
          if (!A_MN) tma_load_2d(&tmA, sa, full_bar(s), k0, m0);
          else
#pragma unroll
            for (int g = 0; g < TBM / 32; ++g) tma_load_2d(&tmA, sa + g * 4096, full_bar(s), m0 + g * 32, k0);
          if (!B_MN) tma_load_2d(&tmB, sb, full_bar(s), k0, nb0);
          else
#pragma unroll
            for (int g = 0; g < BROWS / 32; ++g) tma_load_2d(&tmB, sb + g * 4096, full_bar(s), nb0 + g * 32, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {                                     // ---- MMA issuer
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) {
        if (CTA2) umma_tf32_2sm(d, ad, bd, idesc, acc); else umma_tf32(d, ad, bd, idesc, acc);
      };
      uint32_t it = 0;                                             // k-block counter: stage it % S, chunk buffer it & 1
      for (int t = cid; t < total; t += ncl) {
        int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t cb = it & 1;
          wait_leader(cempty_bar(cb), ((it >> 1) & 1) ^ 1);        // chunk accumulator drained
          wait_leader(conv_bar(s), (it / S) & 1);                  // raw tiles landed and lo tiles written, both CTAs
          if (trace && it < 40) p.dbg[120 + it] = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem_base + cb * BN;
          const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
          uint64_t ad[4], bd[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t oa = A_MN ? sa + k * 1024 : sa + k * 32, ob = B_MN ? sb + k * 1024 : sb + k * 32;
            ad[k] = A_MN ? umma_desc(oa, 4096, 512, 1) : umma_desc(oa, 16, 1024, 2);
            bd[k] = B_MN ? umma_desc(ob, 4096, 512, 1) : umma_desc(ob, 16, 1024, 2);
          }
          constexpr uint64_t LO = (uint64_t)(RAW >> 4);            // the lo tiles sit RAW bytes after the raw ones (16-byte units)
          if (!(dbg & 16)) {
            if (!(dbg & 2)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {                        // cross terms first, into the still tiny accumulator
                mma(tacc, ad[k], bd[k] + LO, k != 0 ? 1u : 0u);    // hi * lo
                mma(tacc, ad[k] + LO, bd[k], 1u);                  // lo * hi
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(tacc, ad[k], bd[k], ((dbg & 2) && k == 0) ? 0u : 1u);   // hi * hi
          }
          commit(empty_bar(s));
          commit(cfull_bar(cb));
        }
      }
    }
  } else if (warp < 10) {                                          // ---- epilogue: warps 2..9
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    EpiWarp w{stg_base + (uint32_t)ew * 4096u, add_bar0 + 8u * ew, 0u, lane, lane >> 3, lane & 7,
              ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.ldc % 4 == 0)};
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * WC);
    const bool drains = BN == 128 || half == 0;                    // narrow tile: the second column half does not exist
    uint32_t it = 0;
    for (int t = cid; t < total; t += ncl) {
      int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
      float a0[32], a1[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t cb = it & 1;
        mbar_wait(cfull_bar(cb), (it >> 1) & 1);
        if (trace && threadIdx.x == 64 && it < 20) p.dbg[160 + it] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!(dbg & 8) && drains) {
          float v[32];                                             // (32 columns at a time: 64 sums + 32 loaded values live)
          tmem_ld32(tlane + cb * BN, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) a0[j] += v[j];
          if (BN == 128) {
            tmem_ld32(tlane + cb * BN + 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) a1[j] += v[j];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) arrive_leader(cempty_bar(cb));
        if (trace && threadIdx.x == 64 && it < 20) p.dbg[180 + it] = clock64();
      }
      if (KSP) {
        // park the partial tile in this CTA's shared memory (stage 0 is free: every MMA that read it has completed):
        // row r at r * 512 B, its 16-byte chunk c at position c ^ (r & 31) — conflict-free for these row-per-lane
        // writes and for the row-per-warp reads of the reduction below
        const int r = q * 32 + lane;
        const uint32_t prow = base + (uint32_t)r * 512u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t c0 = (uint32_t)(half * 16 + j), c1 = c0 + 8u;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((c0 ^ (uint32_t)(r & 31)) << 4)), "f"(a0[4 * j]),
                       "f"(a0[4 * j + 1]), "f"(a0[4 * j + 2]), "f"(a0[4 * j + 3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((c1 ^ (uint32_t)(r & 31)) << 4)), "f"(a1[4 * j]),
                       "f"(a1[4 * j + 1]), "f"(a1[4 * j + 2]), "f"(a1[4 * j + 3]) : "memory");
        }
      } else {
      const int gj = n0 + half * WC;
      if (drains && gj < p.N) epi_cols32(p, &tmC, &tmAdd, a0, gj, m0 + q * 32, w);
      if (BN == 128 && gj + 32 < p.N) epi_cols32(p, &tmC, &tmAdd, a1, gj + 32, m0 + q * 32, w);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  } else {                                                         // ---- converters: warps 10..13
    uint32_t it = 0;
    const int ctid = threadIdx.x - 320;
    for (int t = cid; t < total; t += ncl) {
      int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % S;
        mbar_wait(full_bar(s), (it / S) & 1);
        if (trace && ctid == 0 && it < 40) p.dbg[40 + it] = clock64();
        const uint32_t sa = base + s * STAGE;
        if (!(dbg & 1)) x3_split_stage(sa, sa + RAW, RAW, ctid, false);
        __syncwarp();
        if (lane == 0) arrive_leader(conv_bar(s));
        if (trace && ctid == 0 && it < 40) p.dbg[80 + it] = clock64();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (KSP) {
    __syncwarp();
    cluster_sync_all();                                            // every rank's partial tile is in its shared memory
    int m0, n0, kbeg, nkb; tile_coords(cid, m0, n0, kbeg, nkb);
    const int rp = TBM / (int)ks_n;                                // rows of the tile this CTA reduces and stores
    const bool vec_c = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.ldc % 4 == 0);
    const bool vec_add = p.add && ((reinterpret_cast<uintptr_t>(p.add) & 15) == 0) && (p.ldadd % 4 == 0);
    for (int i = threadIdx.x; i < rp * 32; i += Cfg::THREADS) {
      const int r = (int)ks_rank * rp + (i >> 5), c = i & 31;
      const int gi = m0 + r, gj = n0 + 4 * c;
      if (gi >= p.M || gj >= p.N) continue;
      const uint32_t local = base + (uint32_t)r * 512u + (((uint32_t)c ^ (uint32_t)(r & 31)) << 4);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      for (uint32_t z = 0; z < ks_n; ++z) {                        // rank order: the sum does not depend on timing
        uint32_t ra; float x0, x1, x2, x3;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(z));
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(ra) : "memory");
        v[0] += x0; v[1] += x1; v[2] += x2; v[3] += x3;
      }
      const int nv = min(4, p.N - gj);
      if (p.add) {
        const float* ar = p.add + (int64_t)gi * p.ldadd + gj;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec_add && nv == 4) { const float4 a4 = __ldg(reinterpret_cast<const float4*>(ar)); a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w; }
        else for (int e = 0; e < nv; ++e) a[e] = __ldg(ar + e);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = p.act == ACT_GATE ? (a[e] > 0.f ? v[e] : 0.f) : v[e] + a[e];
      }
      if (p.bias) for (int e = 0; e < nv; ++e) v[e] += __ldg(p.bias + gj + e);
      if (p.act != ACT_NONE && p.act != ACT_GATE) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = tc_act(v[e], p.act);
      }
      float* cr = p.C + (int64_t)(p.c_idx ? p.c_idx[gi] : gi) * p.ldc + gj;
      if (p.accum == ACC_ATOMIC) { for (int e = 0; e < nv; ++e) atomicAdd(cr + e, v[e]); }
      else if (vec_c && nv == 4) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (p.accum == ACC_ADD) { const float4 c4 = *reinterpret_cast<const float4*>(cr); o.x += c4.x; o.y += c4.y; o.z += c4.z; o.w += c4.w; }
        *reinterpret_cast<float4*>(cr) = o;
      } else {
        for (int e = 0; e < nv; ++e) cr[e] = p.accum == ACC_ADD ? cr[e] + v[e] : v[e];
      }
    }
    cluster_sync_all();                                            // no rank leaves while its partial tile may still be read
  } else if (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// =============================================================================================
// "x3w": the chunked 3xTF32 scheme on the WIDE pair tile (256 x 256, cta_group::2) for the large products.
// k_tc_gemm_x3 above pays for its 128-column tiles: half the flops per staged byte, and the main loop of this path is
// bound by the latency of the load -> convert -> MMA -> release loop times the bytes that fit in shared memory
// (measured: 143 TFLOP/s against 224 for the single-accumulator 256 x 256 kernel on the same product).  Here:
//   * a chunk is ONE k-block (32 k): its 8 cross-term MMAs (hi*lo, lo*hi) are issued FIRST into the fresh chunk
//     accumulator, while it only holds values 2^-11 of the final size (their truncation is then negligible), followed
//     by the 4 hi*hi MMAs — four truncating accumulations per chunk, no separate cross accumulator;
//   * the two chunk accumulators take the whole TMEM (2 x 256 columns); the 8 epilogue warps drain each finished
//     chunk into 128 FP32 running sums per thread (the kernel runs 384 threads so that each may hold 168 registers).
// Warps: 0 TMA, 1 MMA issue, 2-3 converters, 4-11 epilogue (lane quadrant = warp % 4, column half = (warp-4) / 4).
// =============================================================================================
struct X3wCfg {
  static constexpr int BN = 256, S = 3;
  static constexpr int RAW = A_BYTES + 128 * TBK * 4;              // 32 KB per CTA per stage
  static constexpr int STAGE = 2 * RAW;
  static constexpr int THREADS = 384;
  static constexpr int SMEM = S * STAGE + 1024 + 8 * 4096 + 256;
};

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(X3wCfg::THREADS, 1)
k_tc_gemm_x3w(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAdd, const TcParams p) {
  using Cfg = X3wCfg;
  constexpr int BN = Cfg::BN, S = Cfg::S, RAW = Cfg::RAW, STAGE = Cfg::STAGE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = base + S * STAGE;
  const uint32_t bars = stg_base + 8 * 4096;
  // full[S] empty[S] conv[S] | cfull[2] cempty[2] | tmem slot | add[8]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto conv_bar = [&](int s) { return bars + 8u * (2 * S + s); };
  auto cfull_bar = [&](int b) { return bars + 8u * (3 * S + b); };
  auto cempty_bar = [&](int b) { return bars + 8u * (3 * S + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (3 * S + 4);
  const uint32_t add_bar0 = bars + 8u * (3 * S + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int gm = (p.M + 255) / 256, gn = (p.N + BN - 1) / BN;
  const int splits = (p.K + p.k_chunk - 1) / p.k_chunk;
  const int total = gm * gn * splits;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(conv_bar(s), 4); }   // 2 warps x 2 CTAs
    for (int b = 0; b < 2; ++b) { mbar_init(cfull_bar(b), 1); mbar_init(cempty_bar(b), 16); }                           // 8 warps x 2 CTAs
    for (int w = 0; w < 8; ++w) mbar_init(add_bar0 + 8u * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                                      // set-up above overlapped the previous kernel; its results are visible from here on

  auto tile_coords = [&](int t, int& m0, int& n0, int& kbeg, int& nkb) {
    const int mt = p.n_fast ? (t / gn) % gm : t % gm, nt = p.n_fast ? t % gn : (t / gm) % gn, z = t / (gm * gn);
    m0 = mt * 256 + (int)rank * TBM; n0 = nt * BN; kbeg = z * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    nkb = (kend - kbeg + TBK - 1) / TBK;
  };

  if (warp < 4) {
    if (warp == 0) {
      if (lane == 0) {                                             // ---- TMA producer (bytes land on this CTA's own barrier)
        uint32_t it = 0;
        for (int t = cid; t < total; t += ncl) {
          int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
          const int nb0 = n0 + (int)rank * 128;
          // the operands stream from HBM and the clusters that share a panel request it at the same moment, so a plain
          // load sees DRAM latency: prefetch this cluster's NEXT tile into L2 while the current one is loaded
          int pm0 = 0, pn0 = 0, pk = 0, pnkb = 0;
          const bool pf = p.prefetch && t + ncl < total;
          if (pf) tile_coords(t + ncl, pm0, pn0, pk, pnkb);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % S;
            mbar_wait_cluster(empty_bar(s), ((it / S) & 1) ^ 1);
            mbar_expect_tx(full_bar(s), RAW);
            const int k0 = kbeg + kb * TBK;
            if (pf && kb < pnkb) {
              const int q0 = pk + kb * TBK, qn = pn0 + (int)rank * 128;
              if (!A_MN) tma_prefetch_2d(&tmA, q0, pm0);
              else
#pragma unroll
                for (int g = 0; g < 4; ++g) tma_prefetch_2d(&tmA, pm0 + g * 32, q0);
              if (!B_MN) tma_prefetch_2d(&tmB, q0, qn);
              else
#pragma unroll
                for (int g = 0; g < 4; ++g) tma_prefetch_2d(&tmB, qn + g * 32, q0);
            }
            const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
            if (!A_MN) tma_load_2d(&tmA, sa, full_bar(s), k0, m0);
            else
#pragma unroll
              for (int g = 0; g < 4; ++g) tma_load_2d(&tmA, sa + g * 4096, full_bar(s), m0 + g * 32, k0);
            if (!B_MN) tma_load_2d(&tmB, sb, full_bar(s), k0, nb0);
            else
#pragma unroll
              for (int g = 0; g < 4; ++g) tma_load_2d(&tmB, sb + g * 4096, full_bar(s), nb0 + g * 32, k0);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && leader) {                                   // ---- MMA issuer: one chunk accumulator per k-block
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        uint32_t it = 0;
        for (int t = cid; t < total; t += ncl) {
          int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t cb = it & 1;
            mbar_wait_cluster(cempty_bar(cb), ((it >> 1) & 1) ^ 1); // both CTAs have drained this chunk accumulator
            mbar_wait_cluster(conv_bar(s), (it / S) & 1);          // raw tiles landed and lo tiles written, both CTAs
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + cb * BN;
            const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
            uint64_t ad[4], bd[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t oa = A_MN ? sa + k * 1024 : sa + k * 32, ob = B_MN ? sb + k * 1024 : sb + k * 32;
              ad[k] = A_MN ? umma_desc(oa, 4096, 512, 1) : umma_desc(oa, 16, 1024, 2);
              bd[k] = B_MN ? umma_desc(ob, 4096, 512, 1) : umma_desc(ob, 16, 1024, 2);
            }
            constexpr uint64_t LO = (uint64_t)(RAW >> 4);          // the lo tiles sit RAW bytes after the raw ones (address field, 16-byte units)
#pragma unroll
            for (int k = 0; k < 4; ++k) {                          // cross terms first, into the still tiny accumulator
              umma_tf32_2sm(tacc, ad[k], bd[k] + LO, idesc, k != 0 ? 1u : 0u);   // hi * lo
              umma_tf32_2sm(tacc, ad[k] + LO, bd[k], idesc, 1u);                 // lo * hi
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32_2sm(tacc, ad[k], bd[k], idesc, 1u);   // hi * hi
            umma_commit_2sm(empty_bar(s));
            umma_commit_2sm(cfull_bar(cb));
          }
        }
      }
    } else {                                                       // ---- converters: warps 2-3 of both CTAs
      uint32_t it = 0;
      const int ctid = threadIdx.x - 64;
      for (int t = cid; t < total; t += ncl) {
        int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(full_bar(s), (it / S) & 1);
          const uint32_t sa = base + s * STAGE;
          x3_split_stage_t<64>(sa, sa + RAW, RAW, ctid);
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(conv_bar(s));
        }
      }
    }
  } else {                                                         // ---- epilogue: warps 4..11
    const int ew = warp - 4, q = warp & 3, half = ew >> 2;
    EpiWarp w{stg_base + (uint32_t)ew * 4096u, add_bar0 + 8u * ew, 0u, lane, lane >> 3, lane & 7,
              ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.ldc % 4 == 0)};
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
    uint32_t it = 0;
    for (int t = cid; t < total; t += ncl) {
      int m0, n0, kbeg, nkb; tile_coords(t, m0, n0, kbeg, nkb);
      float a0[32], a1[32], a2[32], a3[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { a0[j] = 0.f; a1[j] = 0.f; a2[j] = 0.f; a3[j] = 0.f; }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t cb = it & 1;
        mbar_wait(cfull_bar(cb), (it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          float v[32];
          tmem_ld32(tlane + cb * BN, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) a0[j] += v[j];
          tmem_ld32(tlane + cb * BN + 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) a1[j] += v[j];
          tmem_ld32(tlane + cb * BN + 64, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) a2[j] += v[j];
          tmem_ld32(tlane + cb * BN + 96, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) a3[j] += v[j];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(cempty_bar(cb));
      }
      const int gj = n0 + half * 128;
      if (gj < p.N) epi_cols32(p, &tmC, &tmAdd, a0, gj, m0 + q * 32, w);
      if (gj + 32 < p.N) epi_cols32(p, &tmC, &tmAdd, a1, gj + 32, m0 + q * 32, w);
      if (gj + 64 < p.N) epi_cols32(p, &tmC, &tmAdd, a2, gj + 64, m0 + q * 32, w);
      if (gj + 96 < p.N) epi_cols32(p, &tmC, &tmAdd, a3, gj + 96, m0 + q * 32, w);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
inline CUtensorMapL2promotion l2_promo() {
  static const char* e = getenv("DX_TC_L2PROMO");
  if (e && e[0] == '2') return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (e && e[0] == '0') return CU_TENSOR_MAP_L2_PROMOTION_NONE;
  return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
}
// 2-D fp32 tensor [rows][cols] with row pitch ld (floats); box = box_cols x box_rows
bool make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
              bool mn_major, bool plain_f32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  // A tensor map is a pure function of (address, extents, pitch, box, flags): a training step re-issues the same few
  // hundred products every iteration, so the encoded maps are kept in a small direct-mapped per-thread table (4-6 driver
  // calls per product launch otherwise: host time that matters when a step is ~850 short dependent launches).
  struct Slot { const float* ptr; int64_t rows, cols, ld; uint32_t box, flags; bool used; CUtensorMap map; };
  constexpr int NSLOT = 2048;
  static thread_local Slot* table = nullptr;
  if (!table) table = static_cast<Slot*>(calloc(NSLOT, sizeof(Slot)));
  const uint32_t boxk = ((uint32_t)box_cols << 16) | (uint32_t)box_rows;
  const uint32_t flags = (mn_major ? 1u : 0u) | (plain_f32 ? 2u : 0u);
  Slot* sl = nullptr;
  if (table) {
    uint64_t h = reinterpret_cast<uintptr_t>(ptr) * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)rows * 0xC2B2AE3D27D4EB4Full) ^ ((uint64_t)cols << 21) ^ ((uint64_t)ld << 42) ^ ((uint64_t)boxk << 7) ^ flags;
    h ^= h >> 29;
    sl = table + (h & (NSLOT - 1));
    if (sl->used && sl->ptr == ptr && sl->rows == rows && sl->cols == cols && sl->ld == ld && sl->box == boxk && sl->flags == flags) {
      *m = sl->map;
      return true;
    }
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const bool ok = enc(m, plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
             l2_promo(),
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  if (ok && sl) { sl->ptr = ptr; sl->rows = rows; sl->cols = cols; sl->ld = ld; sl->box = boxk; sl->flags = flags; sl->map = *m; sl->used = true; }
  return ok;
}

// 2-D bf16 tensor [rows][cols] with row pitch ld (elements); box = 64 columns (128 bytes) x box_rows, 128-byte swizzle
bool make_map_bf16(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, l2_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  static int cache[64] = {0};
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}
// cudaFuncSetAttribute is per device: remember which devices a kernel family was configured on
struct AttrOnce {
  unsigned long long done = 0;
  template <class F> void operator()(F set) {
    int dev = 0; cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done & bit)) { set(); done |= bit; }
  }
};
inline int tc_prefetch() { static const int v = getenv("DX_TC_NO_PREFETCH") ? 0 : 1; return v; }
inline int x3_dbg() { static const int v = [] { const char* e = getenv("DX_X3_DBG"); return e ? atoi(e) : 0; }(); return v; }

template <int BN>
bool launch_tc(dx_stream_t s, const GemmP& g) {
  using Cfg = TcCfg<BN>;
  CUtensorMap ta, tb;
  // K-major: memory [MN rows][reduction cols]; MN-major: memory [reduction rows][MN cols].  TFLOAT32 maps round on load.
  if (g.a_kc) { if (!make_map(&ta, g.A, g.M, g.K, g.lda, TBK, TBM, false)) return false; }
  else        { if (!make_map(&ta, g.A, g.K, g.M, g.lda, 32, TBK, true)) return false; }
  if (g.b_kc) { if (!make_map(&tb, g.B, g.N, g.K, g.ldb, TBK, BN, false)) return false; }
  else        { if (!make_map(&tb, g.B, g.K, g.N, g.ldb, 32, TBK, true)) return false; }
  // C through TMA when it is a plain strided matrix (no row scatter) with 16-byte aligned rows
  const bool tma_store = !g.c_idx && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (g.ldc % 4 == 0) &&
                         !getenv("DX_TC_NO_TMA_STORE");
  CUtensorMap tc;
  if (tma_store) { if (!make_map(&tc, g.C, g.M, g.N, g.ldc, 32, 32, false, true)) return false; }
  else tc = ta;
  const bool add_tma = tma_store && g.add && ((reinterpret_cast<uintptr_t>(g.add) & 15) == 0) && (g.ldadd % 4 == 0);
  CUtensorMap tadd;
  if (add_tma) { if (!make_map(&tadd, g.add, g.M, g.N, g.ldadd, 32, 32, false, true)) return false; }
  else tadd = ta;
  const int gm = (g.M + TBM - 1) / TBM, gn = (g.N + BN - 1) / BN;
  const int num_sms = sm_count();
  int splits = 1;
  if (g.accum == ACC_ATOMIC) {
    // split the batch reduction so that tiles x splits just fits two rounds of the persistent grid
    const int tiles = gm * gn;
    const int want = (num_sms * 2) / tiles;                    // floor: never spill into a third round
    const int maxs = (g.K + TBK * 16 - 1) / (TBK * 16);       // >= 512 reduction rows per split
    splits = want < 1 ? 1 : (want > maxs ? maxs : want);
  }
  // Few-row products with a long reduction (the dgrads of a small batch: M = 128, K = 1536..2048, a handful of tiles):
  // split the reduction as well, so that more than a handful of SMs stream the weight matrix.  The partial results go
  // out through TMA reduce-add; a plain store becomes zero-fill + reduce-add.  (No bias / added matrix: they would be
  // applied once per split; the relu gate is a 0/1 factor and distributes over the partial sums.)
  int accum = g.accum;
  // Forward products (both operands K-major) are never split: the split count depends on the row count, and a patch's
  // inference result must not depend on how many other patches share its batch.
  if (g.accum != ACC_ATOMIC && !(g.a_kc && g.b_kc) && tma_store && gm * gn < 48 && g.K >= 512 && !g.bias && (!g.add || g.act == ACT_GATE) &&
      (g.act == ACT_NONE || g.act == ACT_GATE) && !getenv("DX_TC_NO_SMALL_SPLIT")) {
    const int want = 96 / (gm * gn), maxs = g.K / 256;
    splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (splits > 1) {
      if (g.accum == ACC_STORE) cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)g.N * 4, (size_t)g.M, s);
      accum = ACC_ADD;
    }
  }
  int k_chunk = (g.K + splits - 1) / splits;
  k_chunk = (k_chunk + TBK - 1) / TBK * TBK;
  splits = (g.K + k_chunk - 1) / k_chunk;
  static long long* dbg = nullptr;
  static const bool want_dbg = getenv("DX_TC_DEBUG") != nullptr;
  if (want_dbg && !dbg) cudaMalloc(&dbg, 256 * sizeof(long long));
  if (want_dbg) cudaMemsetAsync(dbg, 0, 256 * sizeof(long long), s);
  static const bool m_fast = getenv("DX_TC_M_FAST") != nullptr;
  TcParams p{g.M, g.N, g.K, g.C, g.ldc, g.c_idx, g.bias, g.add, g.ldadd, g.act, accum, k_chunk, tma_store ? 1 : 0, add_tma ? 1 : 0,
             m_fast ? 0 : 1, 0, 0, tc_prefetch(), want_dbg ? dbg : nullptr};
  const int total_tiles = gm * gn * splits;
  dim3 grid(total_tiles < num_sms ? total_tiles : num_sms);
  static AttrOnce attr;   // per BN instantiation; the operand-major variants share the footprint
  attr([] {
    cudaFuncSetAttribute(k_tc_gemm<BN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm<BN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm<BN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm<BN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  auto run = [&](auto kern) { launch_k(kern, grid, dim3(Cfg::THREADS), Cfg::SMEM, s, 1, ta, tb, tc, tadd, p); };
  if (g.a_kc && g.b_kc) run(k_tc_gemm<BN, false, false>);
  else if (g.a_kc && !g.b_kc) run(k_tc_gemm<BN, false, true>);
  else if (!g.a_kc && !g.b_kc) run(k_tc_gemm<BN, true, true>);
  else run(k_tc_gemm<BN, true, false>);
  ++g_launches;
  if (want_dbg) {
    long long h[256];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[200];
    fprintf(stderr, "[tc trace] M=%d N=%d K=%d BN=%d splits=%d tiles=%d grid=%d | setup %lld | epilogues:", g.M, g.N, g.K, BN,
            splits, total_tiles, (int)grid.x, h[201] - t0);
    for (int i = 0; i < 8 && h[208 + 2 * i]; ++i) fprintf(stderr, " [%lld..%lld]", h[208 + 2 * i] - t0, h[209 + 2 * i] - t0);
    fprintf(stderr, "\n");
    fprintf(stderr, "  producer(empty ok):");
    for (int i = 0; i < 20 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - t0);
    fprintf(stderr, "\n  mma(full ok):      ");
    for (int i = 0; i < 20 && h[64 + i]; ++i) fprintf(stderr, " %lld", h[64 + i] - t0);
    fprintf(stderr, "\n");
  }
  return true;
}

// 2-CTA launch (BN = 256).  Returns false if not applicable.
bool launch_tc2(dx_stream_t s, const GemmP& g) {
  using Cfg = Tc2Cfg;
  static const bool disabled = getenv("DX_TC_NO_CG2") != nullptr;
  if (disabled) return false;
  const int num_sms = sm_count();
  const int gm2 = (g.M + 255) / 256, gn = (g.N + 255) / 256;
  int splits = 1;
  if (g.accum == ACC_ATOMIC) {
    const int tiles = gm2 * gn;
    const int want = num_sms / tiles;                           // two rounds of the 74 clusters
    const int maxs = (g.K + TBK * 16 - 1) / (TBK * 16);
    splits = want < 1 ? 1 : (want > maxs ? maxs : want);
  }
  int k_chunk = (g.K + splits - 1) / splits;
  k_chunk = (k_chunk + TBK - 1) / TBK * TBK;
  splits = (g.K + k_chunk - 1) / k_chunk;
  const int total = gm2 * gn * splits;
  if (total < 37) return false;                                 // too little work for the pair-tile: 1-CTA kernel
  const bool tma_store = !g.c_idx && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (g.ldc % 4 == 0);
  if (!tma_store) return false;
  if (!g.a_kc && g.b_kc) return false;                          // (no product of the path has this form)
  CUtensorMap ta, tb, tc, tadd;
  if (g.a_kc) { if (!make_map(&ta, g.A, g.M, g.K, g.lda, TBK, TBM, false)) return false; }
  else        { if (!make_map(&ta, g.A, g.K, g.M, g.lda, 32, TBK, true)) return false; }
  if (g.b_kc) { if (!make_map(&tb, g.B, g.N, g.K, g.ldb, TBK, 128, false)) return false; }   // half of the B tile per CTA
  else        { if (!make_map(&tb, g.B, g.K, g.N, g.ldb, 32, TBK, true)) return false; }
  if (!make_map(&tc, g.C, g.M, g.N, g.ldc, 32, 32, false, true)) return false;
  const bool add_tma = g.add && ((reinterpret_cast<uintptr_t>(g.add) & 15) == 0) && (g.ldadd % 4 == 0);
  if (g.add && !add_tma) return false;
  if (add_tma) { if (!make_map(&tadd, g.add, g.M, g.N, g.ldadd, 32, 32, false, true)) return false; }
  else tadd = ta;
  static long long* dbg = nullptr;
  static const bool want_dbg = getenv("DX_TC_DEBUG") != nullptr;
  if (want_dbg && !dbg) cudaMalloc(&dbg, 256 * sizeof(long long));
  if (want_dbg) cudaMemsetAsync(dbg, 0, 256 * sizeof(long long), s);
  static const bool m_fast = getenv("DX_TC_M_FAST") != nullptr;
  TcParams p{g.M, g.N, g.K, g.C, g.ldc, g.c_idx, g.bias, g.add, g.ldadd, g.act, g.accum, k_chunk, 1, add_tma ? 1 : 0,
             m_fast ? 0 : 1, 0, 0, tc_prefetch(), want_dbg ? dbg : nullptr};
  const int ncl = total < num_sms / 2 ? total : num_sms / 2;
  static AttrOnce attr;
  attr([] {
    cudaFuncSetAttribute(k_tc_gemm2<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm2<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm2<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  dim3 grid(2 * ncl);
  if (g.a_kc && g.b_kc) launch_k(k_tc_gemm2<false, false, false>, grid, dim3(Cfg::THREADS), Cfg::SMEM, s, 1, ta, tb, tc, tadd, p);
  else if (g.a_kc && !g.b_kc) launch_k(k_tc_gemm2<false, true, false>, grid, dim3(Cfg::THREADS), Cfg::SMEM, s, 1, ta, tb, tc, tadd, p);
  else launch_k(k_tc_gemm2<true, true, false>, grid, dim3(Cfg::THREADS), Cfg::SMEM, s, 1, ta, tb, tc, tadd, p);
  ++g_launches;
  if (want_dbg) {
    long long h[256];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[200];
    fprintf(stderr, "[tc2 trace] M=%d N=%d K=%d splits=%d pair-tiles=%d clusters=%d | setup %lld | epilogues:", g.M, g.N, g.K,
            splits, total, ncl, h[201] - t0);
    for (int i = 0; i < 8 && h[208 + 2 * i]; ++i) fprintf(stderr, " [%lld..%lld]", h[208 + 2 * i] - t0, h[209 + 2 * i] - t0);
    fprintf(stderr, "\n  mma(full ok):");
    for (int i = 0; i < 20 && h[64 + i]; ++i) fprintf(stderr, " %lld", h[64 + i] - t0);
    fprintf(stderr, "\n");
  }
  return true;
}

// Chunked 3xTF32 launch (k_tc_gemm_x3): CTA2 = pair tiles 256 x 128, else 128 x 128.  Returns false if not applicable.
template <bool CTA2, int BNT = 128>
bool launch_x3(dx_stream_t s, const GemmP& g) {
  using Cfg = X3Cfg<CTA2, BNT>;
  constexpr int TM = CTA2 ? 256 : 128, BN = Cfg::BN;
  const int num_sms = sm_count();
  const int ncta = CTA2 ? num_sms / 2 : num_sms;               // concurrent tiles
  const int gm = (g.M + TM - 1) / TM, gn = (g.N + BN - 1) / BN;
  const bool tma_store = !g.c_idx && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (g.ldc % 4 == 0);
  if (CTA2 && !tma_store) return false;
  int splits = 1, accum = g.accum;
  if (g.accum == ACC_ATOMIC) {
    const int want = (2 * ncta) / (gm * gn);                    // two rounds of the persistent grid
    const int maxs = (g.K + TBK * 16 - 1) / (TBK * 16);        // >= 512 reduction rows per split
    splits = want < 1 ? 1 : (want > maxs ? maxs : want);
  } else if (!CTA2 && !(g.a_kc && g.b_kc) && tma_store && gm * gn < 48 && g.K >= 512 && !g.bias && (!g.add || g.act == ACT_GATE) &&
             (g.act == ACT_NONE || g.act == ACT_GATE)) {
    // few-row dgrads with a long reduction: split it so that more SMs stream the weight matrix (see launch_tc)
    const int want = 96 / (gm * gn), maxs = g.K / 256;
    splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (splits > 1) {
      if (g.accum == ACC_STORE) cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)g.N * 4, (size_t)g.M, s);
      accum = ACC_ADD;
    }
  }
  int k_chunk = (g.K + splits - 1) / splits;
  k_chunk = (k_chunk + TBK - 1) / TBK * TBK;
  splits = (g.K + k_chunk - 1) / k_chunk;
  const int total = gm * gn * splits;
  if (CTA2 && total < ncta / 2) return false;                   // too little work for pair tiles
  CUtensorMap ta, tb, tc, tadd;
  if (g.a_kc) { if (!make_map(&ta, g.A, g.M, g.K, g.lda, TBK, TBM, false, true)) return false; }
  else        { if (!make_map(&ta, g.A, g.K, g.M, g.lda, 32, TBK, true, true)) return false; }
  if (g.b_kc) { if (!make_map(&tb, g.B, g.N, g.K, g.ldb, TBK, Cfg::BROWS, false, true)) return false; }
  else        { if (!make_map(&tb, g.B, g.K, g.N, g.ldb, 32, TBK, true, true)) return false; }
  if (tma_store) { if (!make_map(&tc, g.C, g.M, g.N, g.ldc, 32, 32, false, true)) return false; }
  else tc = ta;
  const bool add_tma = tma_store && g.add && ((reinterpret_cast<uintptr_t>(g.add) & 15) == 0) && (g.ldadd % 4 == 0);
  if (CTA2 && g.add && !add_tma) return false;
  if (add_tma) { if (!make_map(&tadd, g.add, g.M, g.N, g.ldadd, 32, 32, false, true)) return false; }
  else tadd = ta;
  static const bool m_fast = getenv("DX_TC_M_FAST") != nullptr;
  static long long* dbg = nullptr;
  static const bool want_dbg = getenv("DX_TC_DEBUG") != nullptr;
  if (want_dbg && !dbg) cudaMalloc(&dbg, 256 * sizeof(long long));
  if (want_dbg) cudaMemsetAsync(dbg, 0, 256 * sizeof(long long), s);
  TcParams p{g.M, g.N, g.K, g.C, g.ldc, g.c_idx, g.bias, g.add, g.ldadd, g.act, accum, k_chunk, tma_store ? 1 : 0, add_tma ? 1 : 0,
             m_fast ? 0 : 1, x3_dbg(), 0, tc_prefetch(), want_dbg ? dbg : nullptr};
  static AttrOnce attr;
  attr([] {
    cudaFuncSetAttribute(k_tc_gemm_x3<false, false, CTA2, BNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm_x3<false, true, CTA2, BNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm_x3<true, true, CTA2, BNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm_x3<true, false, CTA2, BNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  const int nt = total < ncta ? total : ncta;
  dim3 grid(CTA2 ? 2 * nt : nt);
  auto run = [&](auto kern) {
    launch_k(kern, grid, dim3(Cfg::THREADS), Cfg::SMEM, s, CTA2 ? 2 : 1, ta, tb, tc, tadd, p);
  };
  if (g.a_kc && g.b_kc) run(k_tc_gemm_x3<false, false, CTA2, BNT>);
  else if (g.a_kc && !g.b_kc) run(k_tc_gemm_x3<false, true, CTA2, BNT>);
  else if (!g.a_kc && !g.b_kc) run(k_tc_gemm_x3<true, true, CTA2, BNT>);
  else run(k_tc_gemm_x3<true, false, CTA2, BNT>);
  ++g_launches;
  if (want_dbg) {
    long long h[256];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[250];
    fprintf(stderr, "[x3 trace] M=%d N=%d K=%d cta2=%d splits=%d tiles=%d\n", g.M, g.N, g.K, (int)CTA2, splits, total);
    const char* nm[4] = {"tma issue (empty ok)", "conv start (full ok)", "conv done            ", "mma start (conv ok) "};
    for (int r = 0; r < 4; ++r) {
      fprintf(stderr, "  %s:", nm[r]);
      for (int i = 0; i < 36 && h[40 * r + i]; ++i) fprintf(stderr, " %lld", h[40 * r + i] - t0);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "  chunk full seen by epilogue:");
    for (int i = 0; i < 18 && h[160 + i]; ++i) fprintf(stderr, " %lld", h[160 + i] - t0);
    fprintf(stderr, "\n  chunk drained              :");
    for (int i = 0; i < 18 && h[180 + i]; ++i) fprintf(stderr, " %lld", h[180 + i] - t0);
    fprintf(stderr, "\n");
  }
  return true;
}

// Few-row products (M <= 256: the training step of a small batch, BASELINE config 2) on thread-block clusters:
// k_tc_gemm_x3<.., KSP>.  Measured on B200 (DX_TC_DEBUG clock trace of M = 128, N = 1536, K = 512): a CTA's k-blocks
// are a serial chain of ~1400 clocks each — the 12 MMAs of a 3xTF32 k-block cost ~80-100 clocks apiece whether the tile
// is 32 or 128 columns wide — behind ~5000 clocks of set-up, first TMA, first conversion and the first (slow) MMA
// chunk, so a 16-k-block product takes 14 us on 12 SMs.  Splitting the reduction over the KS CTAs of a cluster makes
// the chain KS times shorter on KS times as many SMs; the partial tiles are summed through distributed shared memory
// in rank order, so the result is deterministic and bias / activation / gate / accumulate epilogues all work (no
// zero fill, no reduce-add).  Forward products take this path only inside the training entry points (FwdSplitScope):
// the split changes the summation order, and inference results must not depend on the batch size.
bool launch_x3k(dx_stream_t s, const GemmP& g) {
  using Cfg = X3Cfg<false, 128>;
  static const bool off = getenv("DX_X3_NO_KSPLIT") != nullptr;
  if (off || g.accum == ACC_ATOMIC || g.M > get_few_rows()) return false;   // (256 rows, 1024 inside a small-batch training step: dx_gemm.h)
  if (g.a_kc && g.b_kc && !get_fwd_split()) return false;
  if (!g.a_kc && g.b_kc) return false;
  const int nkb = (g.K + TBK - 1) / TBK;
  const int tiles = ((g.M + TBM - 1) / TBM) * ((g.N + 127) / 128);
  int ks = 8;
  while (ks > 1 && (tiles * ks > sm_count() || nkb < 2 * ks)) ks >>= 1;   // at least two k-blocks per rank, one wave of CTAs
  if (ks < 2) return false;
  const int k_chunk = (nkb + ks - 1) / ks * TBK;
  CUtensorMap ta, tb;
  if (g.a_kc) { if (!make_map(&ta, g.A, g.M, g.K, g.lda, TBK, TBM, false, true)) return false; }
  else        { if (!make_map(&ta, g.A, g.K, g.M, g.lda, 32, TBK, true, true)) return false; }
  if (g.b_kc) { if (!make_map(&tb, g.B, g.N, g.K, g.ldb, TBK, 128, false, true)) return false; }
  else        { if (!make_map(&tb, g.B, g.K, g.N, g.ldb, 32, TBK, true, true)) return false; }
  TcParams p{g.M, g.N, g.K, g.C, g.ldc, g.c_idx, g.bias, g.add, g.ldadd, g.act, g.accum, k_chunk, 0, 0, 1, 0, 0, 0, nullptr};
  static AttrOnce attr;
  attr([] {
    cudaFuncSetAttribute(k_tc_gemm_x3<false, false, false, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm_x3<false, true, false, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm_x3<true, true, false, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  const dim3 grid(tiles * ks), block(Cfg::THREADS);
  if (g.a_kc && g.b_kc) launch_k(k_tc_gemm_x3<false, false, false, 128, true>, grid, block, Cfg::SMEM, s, ks, ta, tb, ta, ta, p);
  else if (g.a_kc && !g.b_kc) launch_k(k_tc_gemm_x3<false, true, false, 128, true>, grid, block, Cfg::SMEM, s, ks, ta, tb, ta, ta, p);
  else launch_k(k_tc_gemm_x3<true, true, false, 128, true>, grid, block, Cfg::SMEM, s, ks, ta, tb, ta, ta, p);
  ++g_launches;
  return true;
}

// Wide chunked 3xTF32 launch (k_tc_gemm_x3w, pair tiles 256 x 256).  Returns false if not applicable.
// Reduction splits for the persistent pair-tile kernels.  `tiles` output tiles run on `ncl` clusters in waves, and a
// last wave that is nearly empty costs as much as a full one (80 tiles on 74 clusters: 54 % of the machine), while
// every tile pays about 3 k-block times of pipeline fill and epilogue.  Modelled efficiency of s splits:
// occupancy of the waves x kb / (kb + 3) with kb k-blocks per tile; the smallest s within 4 % of the best wins.
inline int choose_splits(int tiles, int K, int ncl, int max_splits) {
  int best = 1; double best_e = 0.0;
  const int kbs = (K + TBK - 1) / TBK;
  for (int sp = 1; sp <= max_splits && sp <= 16; ++sp) {
    const int kb = (kbs + sp - 1) / sp;
    if (sp > 1 && kb < 8) break;                                  // >= 256 reduction rows per split
    const int total = tiles * sp, waves = (total + ncl - 1) / ncl;
    const double e = (double)total / ((double)waves * ncl) * kb / (kb + 3.0);
    if (e > best_e * 1.04) { best_e = e; best = sp; }
  }
  return best;
}

bool launch_x3w(dx_stream_t s, const GemmP& g) {
  using Cfg = X3wCfg;
  static const bool off = getenv("DX_X3_NO_WIDE") != nullptr;
  if (off) return false;
  const int ncl = sm_count() / 2;
  const int gm = (g.M + 255) / 256, gn = (g.N + 255) / 256;
  const bool tma_store = !g.c_idx && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (g.ldc % 4 == 0);
  if (!tma_store || g.N < 192 || (!g.a_kc && g.b_kc)) return false;
  int splits = 1, accum = g.accum;
  static const bool no_model = getenv("DX_X3_NO_SPLIT_MODEL") != nullptr;
  if (g.accum == ACC_ATOMIC) {
    const int want = (2 * ncl) / (gm * gn);
    const int maxs = (g.K + TBK * 16 - 1) / (TBK * 16);
    splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (!no_model) splits = choose_splits(gm * gn, g.K, ncl, maxs);
  } else if (!no_model && !(g.a_kc && g.b_kc) && !g.bias && (!g.add || g.act == ACT_GATE) && (g.act == ACT_NONE || g.act == ACT_GATE)) {
    // dgrads (never the forward products: their value must not depend on the row count): partial sums leave through
    // TMA reduce-add; a plain store becomes zero-fill + reduce-add; the relu gate is a 0/1 factor and distributes
    // (only under one wave of tiles: measured on B200, splitting an 80-tile dgrad over 74 clusters buys nothing — the
    // zero fill and the reduce-adds cost what the better last wave saves — while 16 tiles -> 64 runs 1.56x faster)
    if (gm * gn < ncl) splits = choose_splits(gm * gn, g.K, ncl, g.K / 256);
    if (splits > 1) accum = ACC_ADD;                              // (zero fill just before the launch, below)
  }
  int k_chunk = (g.K + splits - 1) / splits;
  k_chunk = (k_chunk + TBK - 1) / TBK * TBK;
  splits = (g.K + k_chunk - 1) / k_chunk;
  const int total = gm * gn * splits;
  if (total < 37) return false;
  CUtensorMap ta, tb, tc, tadd;
  if (g.a_kc) { if (!make_map(&ta, g.A, g.M, g.K, g.lda, TBK, TBM, false, true)) return false; }
  else        { if (!make_map(&ta, g.A, g.K, g.M, g.lda, 32, TBK, true, true)) return false; }
  if (g.b_kc) { if (!make_map(&tb, g.B, g.N, g.K, g.ldb, TBK, 128, false, true)) return false; }
  else        { if (!make_map(&tb, g.B, g.K, g.N, g.ldb, 32, TBK, true, true)) return false; }
  if (!make_map(&tc, g.C, g.M, g.N, g.ldc, 32, 32, false, true)) return false;
  const bool add_tma = g.add && ((reinterpret_cast<uintptr_t>(g.add) & 15) == 0) && (g.ldadd % 4 == 0);
  if (g.add && !add_tma) return false;
  if (add_tma) { if (!make_map(&tadd, g.add, g.M, g.N, g.ldadd, 32, 32, false, true)) return false; }
  else tadd = ta;
  static const bool m_fast = getenv("DX_TC_M_FAST") != nullptr;
  TcParams p{g.M, g.N, g.K, g.C, g.ldc, g.c_idx, g.bias, g.add, g.ldadd, g.act, accum, k_chunk, 1, add_tma ? 1 : 0,
             m_fast ? 0 : 1, 0, 1, tc_prefetch(), nullptr};
  static AttrOnce attr;
  attr([] {
    cudaFuncSetAttribute(k_tc_gemm_x3w<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm_x3w<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(k_tc_gemm_x3w<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  const int nt = total < ncl ? total : ncl;
  if (accum != g.accum && g.accum == ACC_STORE) cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)g.N * 4, (size_t)g.M, s);
  const dim3 grid(2 * nt), block(Cfg::THREADS);
  if (g.a_kc && g.b_kc) launch_k(k_tc_gemm_x3w<false, false>, grid, block, Cfg::SMEM, s, 2, ta, tb, tc, tadd, p);
  else if (g.a_kc && !g.b_kc) launch_k(k_tc_gemm_x3w<false, true>, grid, block, Cfg::SMEM, s, 2, ta, tb, tc, tadd, p);
  else launch_k(k_tc_gemm_x3w<true, true>, grid, block, Cfg::SMEM, s, 2, ta, tb, tc, tadd, p);
  ++g_launches;
  return true;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// C[M,N] = act(A B^T + bias) with bf16 K-major operands (A [M,K], B [N,K], pitches in elements), fp32 accumulate / output.
bool launch_tc2_bf16(dx_stream_t s, const GemmP& g, const void* A16, const void* B16) {
  if (g.accum != ACC_STORE || g.c_idx || g.add || !al16(A16) || !al16(B16) || (g.lda % 8) || (g.ldb % 8) || !al16(g.C) || (g.ldc % 4))
    return false;
  CUtensorMap ta, tb, tc;
  if (!make_map_bf16(&ta, A16, g.M, g.K, g.lda, TBM) || !make_map_bf16(&tb, B16, g.N, g.K, g.ldb, 128) ||
      !make_map(&tc, g.C, g.M, g.N, g.ldc, 32, 32, false, true))
    return false;
  const int gm2 = (g.M + 255) / 256, gn = (g.N + 255) / 256;
  const int k_chunk = (g.K + 63) / 64 * 64;
  TcParams p{g.M, g.N, g.K, g.C, g.ldc, nullptr, g.bias, nullptr, 0, g.act, ACC_STORE, k_chunk, 1, 0, 1, 0, 0, 0, nullptr};
  const int num_sms = sm_count();
  const int total = gm2 * gn, ncl = total < num_sms / 2 ? total : num_sms / 2;
  static AttrOnce attr;
  attr([] { cudaFuncSetAttribute(k_tc_gemm2<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg::SMEM); });
  launch_k(k_tc_gemm2<false, false, true>, dim3(2 * ncl), dim3(320), Tc2Cfg::SMEM, s, 1, ta, tb, tc, ta, p);
  ++g_launches;
  return true;
}

}  // namespace

bool tc_gemm_bf16(dx_stream_t s, const GemmP& g, const void* A16, const void* B16) { return launch_tc2_bf16(s, g, A16, B16); }

// Returns false when the problem is not eligible (caller falls back to the FP32 SIMT kernel).
// x3: error-compensated 3xTF32 (FP32-accurate) instead of plain TF32.
bool tc_gemm(dx_stream_t s, const GemmP& g, int* tile_n, bool x3) {
  if (g.a_idx || g.b_idx) return false;                       // TMA tiles cannot gather rows
  // TMA needs 16-byte aligned bases and row pitches; everything else (ragged M/N/K) is handled by
  // the tensor maps' out-of-bounds zero fill / clipping.
  if (!al16(g.A) || !al16(g.B) || (g.lda % 4) || (g.ldb % 4)) return false;
  // Products with a tiny weight side go to the FFMA kernels.  The rule looks at N x K only, never at the row count: which
  // kernel family computes a forward product must not depend on how many patches share the batch (a one-graph batch
  // and a 32768-graph batch give a patch the same bits).
  // (Weight gradients — ACC_ATOMIC, K = the batch rows of the step — are exempt: a 512 x 512 gradient over the 2-12 active
  // rows of a small step is 16 short tiles here, but 262 k atomic adds in the FFMA kernel: 34 us against ~10.)
  if (g.accum != ACC_ATOMIC && (double)g.N * g.K < 4096.0) return false;
  if (g.accum == ACC_ATOMIC && (double)g.M * g.N < 4096.0) return false;
  int bn = g.N >= 192 ? 256 : (g.N >= 96 ? 128 : 64);
  // Small batches (M of a few hundred rows): a 128 x 256 tiling leaves most SMs idle and each CTA streams a large
  // slice of the weight matrix alone; narrower tiles spread that stream over more SMs (latency, not throughput).
  auto ntiles = [&](int b) { return ((g.M + TBM - 1) / TBM) * ((g.N + b - 1) / b); };
  if (g.accum != ACC_ATOMIC) {
    if (bn == 256 && ntiles(256) < 48) bn = 128;
    if (bn == 128 && ntiles(128) < 48) bn = 64;
  }
  if (tile_n) *tile_n = bn;
  if (x3) {
    static const bool no_narrow = getenv("DX_X3_NO_NARROW") != nullptr;
    if (launch_x3k(s, g)) return true;                                // few rows, long reduction: cluster split-K
    if (g.N <= 64 && !no_narrow) return launch_x3<false, 32>(s, g);   // thin outputs: 32-column tiles
    // Few-row products (small training batches, M <= 256): a 128-column tiling runs N/128 CTAs whose k-blocks are a serial
    // load -> convert -> MMA -> drain chain; 32-column tiles put 4x as many SMs on the weight stream and shorten every link
    // of the chain.  Same per-element instruction sequence as the other tile shapes, so forward results keep their bits.
    static const int few_rows = [] { const char* e = getenv("DX_X3_FEW_ROWS"); return e ? atoi(e) : 256; }();   // (0 disables)
    if (few_rows > 0 && g.accum != ACC_ATOMIC && g.M <= few_rows && !no_narrow) return launch_x3<false, 32>(s, g);
    // wide pair tiles whenever the output is wide enough: for dgrads / wgrads launch_x3w splits the reduction until the
    // machine is filled (choose_splits), forward products need >= 37 tiles of their own
    if (g.N >= 192 && launch_x3w(s, g)) return true;
    if (launch_x3<true>(s, g)) return true;
    return launch_x3<false>(s, g);
  }
  if (bn == 256 && launch_tc2(s, g)) return true;
  return bn == 256 ? launch_tc<256>(s, g) : (bn == 128 ? launch_tc<128>(s, g) : launch_tc<64>(s, g));
}

}  // namespace dx
#else
namespace dx {
bool tc_gemm(dx_stream_t, const GemmP&, int*, bool) { return false; }
bool tc_gemm_bf16(dx_stream_t, const GemmP&, const void*, const void*) { return false; }
}  // namespace dx
#endif
