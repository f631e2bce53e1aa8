// Shared definitions for the dxvae_b200 native library.
//
// Build modes
//   default   nvcc, sm_100a: kernels run on the GPU (the product).
//   DX_EMU    g++ only, used by tests/emu: the SAME host orchestration and the SAME
//             per-element device functors are executed serially on the CPU so the
//             schedule / buffer / gradient-chain logic can be checked in the build
//             container, which has no GPU.  The product library never contains this
//             mode and the Python package never loads an emulation build.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#ifdef DX_EMU
#define DX_HD
#define DX_D
#define DX_INLINE inline
typedef void* dx_stream_t;
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
#else
#include <cuda_runtime.h>
#define DX_HD __host__ __device__
#define DX_D __device__
#define DX_INLINE __forceinline__
typedef cudaStream_t dx_stream_t;
#endif

namespace dx {

constexpr int NN = 7;      // nodes
constexpr int NP = 21;     // params per node
constexpr int SX = 27;     // node feature width
constexpr int SX0 = 23;    // node-0 feature width
constexpr int H = 512;     // hidden
constexpr int Z = 128;     // latent
constexpr int XP = 32;     // padded feature width (node-major Xn rows are 128 B)
constexpr int G3 = 3 * H;  // GRU gate width (r,z,n)

// ---- parameter table (state_dict order of model.py:24-72) -----------------------
enum ParamId {
  P_CE_WIH, P_CE_WHH, P_CE_BIH, P_CE_BHH,   // combin_encode
  P_LE_WIH, P_LE_WHH, P_LE_BIH, P_LE_BHH,   // loop_encode
  P_RE_WIH, P_RE_WHH, P_RE_BIH, P_RE_BHH,   // root_encode
  P_MU_W, P_MU_B, P_STD_W, P_STD_B,         // h_to_mu, h_to_std.0
  P_CD_WIH, P_CD_WHH, P_CD_BIH, P_CD_BHH,   // combin_decode
  P_LD_WIH, P_LD_WHH, P_LD_BIH, P_LD_BHH,   // loop_decode
  P_RD_WIH, P_RD_WHH, P_RD_BIH, P_RD_BHH,   // root_decode
  P_ZH_W, P_ZH_B,                           // z_to_h.0
  P_X0_W0, P_X0_B0, P_X0_W2, P_X0_B2, P_X0_W4, P_X0_B4,   // h_to_x0.{0,2,4}
  P_X_W0, P_X_B0, P_X_W2, P_X_B2, P_X_W4, P_X_B4,         // h_to_x.{0,2,4}
  P_ES_W0, P_ES_B0, P_ES_W2, P_ES_B2,       // h_to_edge_self.{0,2}
  P_E_W0, P_E_B0, P_E_W2, P_E_B2,           // h_to_edge.{0,2}
  P_G_W, P_G_B,                             // gate.0
  P_M_W,                                    // mapper.0
  P_COUNT
};
static_assert(P_COUNT == 53, "state_dict has 53 tensors (SURVEY App. E lists them; its count of 46 is off)");

struct ParamDesc { const char* name; int rows, cols; };
// 53 tensors; GRU cells contribute 4 each.
static const ParamDesc kParams[P_COUNT] = {
  {"combin_encode.weight_ih", G3, SX}, {"combin_encode.weight_hh", G3, H}, {"combin_encode.bias_ih", G3, 0}, {"combin_encode.bias_hh", G3, 0},
  {"loop_encode.weight_ih", G3, SX}, {"loop_encode.weight_hh", G3, H}, {"loop_encode.bias_ih", G3, 0}, {"loop_encode.bias_hh", G3, 0},
  {"root_encode.weight_ih", G3, SX0}, {"root_encode.weight_hh", G3, H}, {"root_encode.bias_ih", G3, 0}, {"root_encode.bias_hh", G3, 0},
  {"h_to_mu.weight", Z, H}, {"h_to_mu.bias", Z, 0}, {"h_to_std.0.weight", Z, H}, {"h_to_std.0.bias", Z, 0},
  {"combin_decode.weight_ih", G3, SX}, {"combin_decode.weight_hh", G3, H}, {"combin_decode.bias_ih", G3, 0}, {"combin_decode.bias_hh", G3, 0},
  {"loop_decode.weight_ih", G3, SX}, {"loop_decode.weight_hh", G3, H}, {"loop_decode.bias_ih", G3, 0}, {"loop_decode.bias_hh", G3, 0},
  {"root_decode.weight_ih", G3, SX0}, {"root_decode.weight_hh", G3, H}, {"root_decode.bias_ih", G3, 0}, {"root_decode.bias_hh", G3, 0},
  {"z_to_h.0.weight", H, Z}, {"z_to_h.0.bias", H, 0},
  {"h_to_x0.0.weight", 2 * H, H}, {"h_to_x0.0.bias", 2 * H, 0}, {"h_to_x0.2.weight", 2 * H, 2 * H}, {"h_to_x0.2.bias", 2 * H, 0},
  {"h_to_x0.4.weight", SX0 + 32, 2 * H}, {"h_to_x0.4.bias", SX0 + 32, 0},
  {"h_to_x.0.weight", 2 * H, H}, {"h_to_x.0.bias", 2 * H, 0}, {"h_to_x.2.weight", 2 * H, 2 * H}, {"h_to_x.2.bias", 2 * H, 0},
  {"h_to_x.4.weight", SX, 2 * H}, {"h_to_x.4.bias", SX, 0},
  {"h_to_edge_self.0.weight", 2 * H, H}, {"h_to_edge_self.0.bias", 2 * H, 0}, {"h_to_edge_self.2.weight", 1, 2 * H}, {"h_to_edge_self.2.bias", 1, 0},
  {"h_to_edge.0.weight", 4 * H, 2 * H}, {"h_to_edge.0.bias", 4 * H, 0}, {"h_to_edge.2.weight", 2, 4 * H}, {"h_to_edge.2.bias", 2, 0},
  {"gate.0.weight", H, 2 * H}, {"gate.0.bias", H, 0},
  {"mapper.0.weight", H, 2 * H},
};

constexpr int64_t kAlign = 64;  // floats: every tensor starts 256-byte aligned
inline int64_t param_numel(int k) { return (int64_t)kParams[k].rows * (kParams[k].cols ? kParams[k].cols : 1); }
inline int64_t param_offset(int k) {
  int64_t off = 0;
  for (int i = 0; i < k; ++i) off += (param_numel(i) + kAlign - 1) / kAlign * kAlign;
  return off;
}
inline int64_t param_blob_floats() { return param_offset(P_COUNT); }

struct Offsets {
  int64_t o[P_COUNT];
  Offsets() { for (int k = 0; k < P_COUNT; ++k) o[k] = param_offset(k); }
};
inline const Offsets& offsets() { static Offsets s; return s; }

// ---- errors ------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define DX_CHECK(cond, ...) do { if (!(cond)) { ::dx::set_error(__VA_ARGS__); return 1; } } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace dx
