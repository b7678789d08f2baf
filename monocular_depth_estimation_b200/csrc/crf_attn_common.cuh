// Shared pieces of the window-attention kernels: parameter block, the tile-row -> token map, 64-byte head-slice
// gathers into swizzled UMMA tiles, relative-position-bias indexing.
#pragma once
#include "crf_host.h"
#include "crf_ptx.cuh"
#include "crf_window.cuh"

namespace crf {

constexpr int kNTok = 49;

struct AttnParams {
  WindowGeom gm;
  int B, C, nH;
  int total_windows;  // B * nW
  int npairs;
  const __nv_bfloat16* qk;   // (T, 2C)
  const __nv_bfloat16* vb;   // (T, C)
  const float* qk_bias;      // (2C)
  const float* table;        // (169, nH)
  float scale;
  const float* ext_mask;     // optional additive mask (nwm, 49, 49) f32, window index = global window % nwm; or nullptr
  int ext_mask_nw;
  // forward
  __nv_bfloat16* o;          // (T, C)
  float* lse;                // (B*nW, nH, 64)
  // backward
  const __nv_bfloat16* dout; // (T, C)
  __nv_bfloat16* dqk;        // (T, 2C)
  float* dv;                 // (T, C)
  int dv_acc;
  float* d_table;            // (169, nH)
  float* d_qk_bias;          // (2C)
  float rcp_nW, rcp_nWw;     // 1 / windows per image, 1 / windows per row (fast exact division in the kernels)
  int prof;                  // development: block (0,0) prints per-phase cycle counts (CRF_ATTN_PROF=1)
};

// Token index of tile row r of a window pair: >= 0 real token, -1 zero-pad token, -2 dead row.
__device__ __forceinline__ int row_token(const AttnParams& P, int pair, int r, int& window_global, int& pos) {
  const int half = r >> 6;
  pos = r & 63;
  window_global = 2 * pair + half;
  if (pos >= kNTok || window_global >= P.total_windows) return -2;
  const int b = window_global / P.gm.nW;
  const int win = window_global - b * P.gm.nW;
  const int src = P.gm.source(win, pos);
  return src < 0 ? -1 : b * P.gm.H * P.gm.W + src;
}

// copy one 64-byte head slice (32 bf16) of a token row into row r of a SW64 tile
__device__ __forceinline__ void gather_row64(uint32_t tile, int r, const __nv_bfloat16* src) {
#pragma unroll
  for (int c = 0; c < 4; ++c) cp_async16(tile + sw64_offset(r, c), src + 8 * c);
}
__device__ __forceinline__ void zero_row64(uint8_t* tile_gen, int r) {
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(tile_gen + sw64_offset(r, c)) = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void bias_row64(uint8_t* tile_gen, int r, const float* bias32) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias32 + 8 * c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias32 + 8 * c + 4));
    *reinterpret_cast<uint4*>(tile_gen + sw64_offset(r, c)) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

// relative_position_index[i][j] = (yi - yj + 6) * 13 + (xi - xj + 6)   (newcrf_layers.py:90-99)
__device__ __forceinline__ int rpb_base(int pos) { return (pos / 7) * 13 + (pos % 7) + 84; }
__host__ __device__ constexpr int rpb_col(int j) { return (j / 7) * 13 + (j % 7); }

// register re-balancing between warp-specialised roles (whole warpgroups execute these)
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

int fill_attn_params(AttnParams& P, const crf_block_desc& d);

}  // namespace crf
