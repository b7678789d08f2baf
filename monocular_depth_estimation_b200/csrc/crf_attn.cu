// Window-attention core of the CRF block on tcgen05, forward and backward.
//
// Replaces, without materialising any of them in HBM: F.pad, torch.roll, window_partition of x and v, the
// relative-position-bias gather, the shifted-window mask, softmax, attn @ v, window_reverse, the reverse roll and
// the crop (newcrf_layers.py:121-146, :212-249, :332-350).
//
// Work decomposition: a CTA owns one head and loops over PAIRS of windows.  The two 49-token windows of a pair are
// stacked into one 128-row UMMA tile (rows 0..48 = window A, 64..112 = window B, the rest zero):
//   S  [128 x 128] = Q[128 x hd] * [K_A ; K_B]^T        -- row r only uses the 64 columns of its own window
//   O  [128 x hd]  = P[128 x 64] * V_A  |  P * V_B       -- two N=hd MMAs, row r reads the half of its window
// Token rows are gathered straight from the token-major q/k/v tensors with the closed-form pad+roll index map
// (cp.async 16-byte copies into the swizzled UMMA layout); zero-padded tokens get k = bias, v = 0 exactly as the
// reference's pad-after-LayerNorm produces.  Softmax runs one thread per accumulator row (TMEM lane), fp32.
//
// Backward recomputes S and P from q, k and the saved row log-sum-exp, and uses block-diagonal 128 x 128 P / dS
// tiles so every product is a single M=128 accumulation:
//   dP = dO * [V_A;V_B]^T   dV = Pbd^T * dO   dK = dSbd^T * Q   dQ = dSbd * [K_A;K_B]
// The relative-position-bias gradient is accumulated in registers across all pairs a CTA processes (each thread
// owns one query row) and flushed once per CTA.
#include <stdlib.h>

#include "crf_attn_common.cuh"

namespace crf {

namespace {

constexpr int kAttnThreads = 160;  // warps 0-3: one thread per tile row; warp 4: MMA issuer / TMEM owner
// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const AttnParams P) {
  constexpr int HD = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t Qs = base, Ks = base + 8192, Vs = base + 16384, Ps = base + 24576;
  uint8_t* Ks_gen = gen + 8192;
  uint8_t* Vs_gen = gen + 16384;
  uint8_t* Qs_gen = gen;
  uint8_t* Ps_gen = gen + 24576;
  float* tbl = reinterpret_cast<float*>(gen + 40960);                 // 176 floats
  uint8_t* rid = gen + 40960 + 176 * 4;                               // 128 bytes
  const uint32_t bar_s = base + 40960 + 176 * 4 + 128;
  const uint32_t bar_o = bar_s + 8;
  const uint32_t tmem_ptr_addr = bar_s + 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + 40960 + 176 * 4 + 128 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int C = P.C;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_addr, 128);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 169; i += kAttnThreads) tbl[i] = __ldg(P.table + i * P.nH + h);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;

  const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128, 128);
  const uint32_t idesc_o = make_idesc(1u, 0u, 1u, 128, HD);

  int it = 0;
  for (int pair = blockIdx.x; pair < P.npairs; pair += gridDim.x, ++it) {
    int tok = -2, wg = 0, pos = 0;
    const int r = threadIdx.x;
    // ---- phase 0: gather Q, K, V rows of this head ----
    if (warp < 4) {
      tok = row_token(P, pair, r, wg, pos);
      if (tok >= 0) {
        const __nv_bfloat16* qrow = P.qk + static_cast<int64_t>(tok) * 2 * C + h * HD;
        gather_row64(Qs, r, qrow);
        gather_row64(Ks, r, qrow + C);
        gather_row64(Vs, r, P.vb + static_cast<int64_t>(tok) * C + h * HD);
      } else {
        zero_row64(Qs_gen, r);
        zero_row64(Vs_gen, r);
        if (tok == -1) bias_row64(Ks_gen, r, P.qk_bias + C + h * HD);
        else zero_row64(Ks_gen, r);
      }
      int region = 0;
      if (tok != -2 && P.gm.shift > 0) {
        const int b = wg / P.gm.nW;
        region = P.gm.region(wg - b * P.gm.nW, pos);
      }
      rid[r] = static_cast<uint8_t>(region);
      cp_async_commit();
      cp_async_wait_all();
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    // ---- phase 1: S = Q K^T ----
    if (warp == 4 && lane == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        umma_bf16(tmem, make_smem_desc(Qs + ks * 32, 16, 512, kSwizzle64), make_smem_desc(Ks + ks * 32, 16, 512, kSwizzle64),
                  idesc_s, ks > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    // ---- phase 2: softmax, one thread per row ----
    if (warp < 4) {
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      const int half = r >> 6;
      const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + half * 64;
      uint32_t s0[32], s1[32];
      tmem_ld32(taddr, s0);
      tmem_ld32(taddr + 32, s1);
      tmem_ld_wait();
      float p[64];
      if (tok != -2) {
        const int bi = rpb_base(pos);
        const uint8_t* rrow = rid + half * 64;
        const int my_region = rrow[pos];
        const bool masked = P.gm.shift > 0;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kNTok; ++j) {
          float s = __uint_as_float(j < 32 ? s0[j & 31] : s1[j & 31]) + tbl[bi - rpb_col(j)];
          if (masked && rrow[j] != my_region) s += -100.0f;
          if (P.ext_mask != nullptr)
            s += __ldg(P.ext_mask + (static_cast<int64_t>(wg % P.ext_mask_nw) * kNTok + pos) * kNTok + j);
          p[j] = s;
          mx = fmaxf(mx, s);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kNTok; ++j) {
          p[j] = __expf(p[j] - mx);
          sum += p[j];
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int j = 0; j < kNTok; ++j) p[j] *= inv;
#pragma unroll
        for (int j = kNTok; j < 64; ++j) p[j] = 0.f;
        if (P.lse != nullptr) P.lse[(static_cast<int64_t>(wg) * P.nH + h) * 64 + pos] = mx + __logf(sum);
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) p[j] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(Ps_gen + sw128_offset(r, c)) =
            make_uint4(pack_bf16(p[8 * c], p[8 * c + 1]), pack_bf16(p[8 * c + 2], p[8 * c + 3]),
                       pack_bf16(p[8 * c + 4], p[8 * c + 5]), pack_bf16(p[8 * c + 6], p[8 * c + 7]));
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    // ---- phase 3: O = P V (window A -> cols [0,hd), window B -> cols [hd,2hd)) ----
    if (warp == 4 && lane == 0) {
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem + half * HD, make_smem_desc(Ps + ks * 32, 16, 1024, kSwizzle128),
                    make_smem_desc(Vs + half * 4096 + ks * 1024, 512, 512, kSwizzle64), idesc_o, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
    // ---- phase 4: store O rows in token order (window_reverse + un-roll + crop) ----
    if (warp < 4) {
      mbar_wait(bar_o, it & 1);
      tc_fence_after();
      const int half = r >> 6;
      uint32_t o[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + half * HD, o);
      tmem_ld_wait();
      if (tok >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(P.o + static_cast<int64_t>(tok) * C + h * HD);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(pack_bf16(__uint_as_float(o[8 * c]), __uint_as_float(o[8 * c + 1])),
                              pack_bf16(__uint_as_float(o[8 * c + 2]), __uint_as_float(o[8 * c + 3])),
                              pack_bf16(__uint_as_float(o[8 * c + 4]), __uint_as_float(o[8 * c + 5])),
                              pack_bf16(__uint_as_float(o[8 * c + 6]), __uint_as_float(o[8 * c + 7])));
      }
    }
    // the next iteration's first __syncthreads (preceded by tcgen05.fence::before_thread_sync) orders these TMEM
    // reads and the smem reads of the finished MMAs before they are overwritten.
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
  (void)Qs_gen;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_kernel(const AttnParams P) {
  constexpr int HD = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // Q, K, V, dO: 8 KB each (SW64); Pbd, dSbd: 2 x 16 KB each (two SW128 column chunks of 128 rows)
  const uint32_t Qs = base, Ks = base + 8192, Vs = base + 16384, Gs = base + 24576;
  const uint32_t Pb = base + 32768, Db = base + 65536;
  constexpr int kMisc = 98304;
  float* tbl = reinterpret_cast<float*>(gen + kMisc);  // 176 floats
  uint8_t* rid = gen + kMisc + 176 * 4;                // 128 bytes
  const uint32_t bar_s = base + kMisc + 176 * 4 + 128;
  const uint32_t bar_o = bar_s + 8;
  const uint32_t tmem_ptr_addr = bar_s + 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + kMisc + 176 * 4 + 128 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int C = P.C;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_addr, 256);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 169; i += kAttnThreads) tbl[i] = __ldg(P.table + i * P.nH + h);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;

  const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128, 128);   // S, dP : K-major x K-major
  const uint32_t idesc_t = make_idesc(1u, 1u, 1u, 128, HD);    // dV, dK: MN-major x MN-major
  const uint32_t idesc_q = make_idesc(1u, 0u, 1u, 128, HD);    // dQ    : K-major x MN-major

  float dtab[kNTok];  // sum over all my pairs of dS[my row][j]
#pragma unroll
  for (int j = 0; j < kNTok; ++j) dtab[j] = 0.f;
  const int my_pos = threadIdx.x & 63;

  int it = 0;
  for (int pair = blockIdx.x; pair < P.npairs; pair += gridDim.x, ++it) {
    int tok = -2, wg = 0, pos = 0;
    const int r = threadIdx.x;
    if (warp < 4) {
      tok = row_token(P, pair, r, wg, pos);
      if (tok >= 0) {
        const __nv_bfloat16* qrow = P.qk + static_cast<int64_t>(tok) * 2 * C + h * HD;
        gather_row64(Qs, r, qrow);
        gather_row64(Ks, r, qrow + C);
        gather_row64(Vs, r, P.vb + static_cast<int64_t>(tok) * C + h * HD);
        gather_row64(Gs, r, P.dout + static_cast<int64_t>(tok) * C + h * HD);
      } else {
        zero_row64(gen, r);
        zero_row64(gen + 16384, r);
        zero_row64(gen + 24576, r);
        if (tok == -1) bias_row64(gen + 8192, r, P.qk_bias + C + h * HD);
        else zero_row64(gen + 8192, r);
      }
      int region = 0;
      if (tok != -2 && P.gm.shift > 0) {
        const int b = wg / P.gm.nW;
        region = P.gm.region(wg - b * P.gm.nW, pos);
      }
      rid[r] = static_cast<uint8_t>(region);
      cp_async_commit();
      cp_async_wait_all();
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    // ---- S = Q K^T -> cols [0,128);  dP = dO V^T -> cols [128,256) ----
    if (warp == 4 && lane == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        umma_bf16(tmem, make_smem_desc(Qs + ks * 32, 16, 512, kSwizzle64), make_smem_desc(Ks + ks * 32, 16, 512, kSwizzle64),
                  idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        umma_bf16(tmem + 128, make_smem_desc(Gs + ks * 32, 16, 512, kSwizzle64),
                  make_smem_desc(Vs + ks * 32, 16, 512, kSwizzle64), idesc_s, ks > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    // ---- P, dS per row; block-diagonal bf16 tiles ----
    if (warp < 4) {
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      const int half = r >> 6;
      const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + half * 64;
      float p[64], ds[64];
      {
        uint32_t s0[32], s1[32];
        tmem_ld32(taddr, s0);
        tmem_ld32(taddr + 32, s1);
        tmem_ld_wait();
        if (tok >= 0) {
          const float lse = __ldg(P.lse + (static_cast<int64_t>(wg) * P.nH + h) * 64 + pos);
          const int bi = rpb_base(pos);
          const uint8_t* rrow = rid + half * 64;
          const int my_region = rrow[pos];
          const bool masked = P.gm.shift > 0;
#pragma unroll
          for (int j = 0; j < kNTok; ++j) {
            float s = __uint_as_float(j < 32 ? s0[j & 31] : s1[j & 31]) + tbl[bi - rpb_col(j)];
            if (masked && rrow[j] != my_region) s += -100.0f;
            if (P.ext_mask != nullptr)
              s += __ldg(P.ext_mask + (static_cast<int64_t>(wg % P.ext_mask_nw) * kNTok + pos) * kNTok + j);
            p[j] = __expf(s - lse);
          }
        } else {
#pragma unroll
          for (int j = 0; j < kNTok; ++j) p[j] = 0.f;
        }
#pragma unroll
        for (int j = kNTok; j < 64; ++j) p[j] = 0.f;
      }
      {
        uint32_t g0[32], g1[32];
        tmem_ld32(taddr + 128, g0);
        tmem_ld32(taddr + 128 + 32, g1);
        tmem_ld_wait();
        float dsum = 0.f;
#pragma unroll
        for (int j = 0; j < kNTok; ++j) {
          ds[j] = __uint_as_float(j < 32 ? g0[j & 31] : g1[j & 31]);
          dsum += p[j] * ds[j];
        }
#pragma unroll
        for (int j = 0; j < kNTok; ++j) {
          ds[j] = p[j] * (ds[j] - dsum);
          dtab[j] += ds[j];
        }
#pragma unroll
        for (int j = kNTok; j < 64; ++j) ds[j] = 0.f;
      }
      // row r of the 128 x 128 block-diagonal tiles: own window's 64 columns carry data, the other 64 are zero
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t own = static_cast<uint32_t>(half) * 16384u + sw128_offset(r, c);
        const uint32_t oth = static_cast<uint32_t>(half ^ 1) * 16384u + sw128_offset(r, c);
        *reinterpret_cast<uint4*>(gen + 32768 + own) =
            make_uint4(pack_bf16(p[8 * c], p[8 * c + 1]), pack_bf16(p[8 * c + 2], p[8 * c + 3]),
                       pack_bf16(p[8 * c + 4], p[8 * c + 5]), pack_bf16(p[8 * c + 6], p[8 * c + 7]));
        *reinterpret_cast<uint4*>(gen + 32768 + oth) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(gen + 65536 + own) =
            make_uint4(pack_bf16(ds[8 * c], ds[8 * c + 1]), pack_bf16(ds[8 * c + 2], ds[8 * c + 3]),
                       pack_bf16(ds[8 * c + 4], ds[8 * c + 5]), pack_bf16(ds[8 * c + 6], ds[8 * c + 7]));
        *reinterpret_cast<uint4*>(gen + 65536 + oth) = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    // ---- dV -> cols [0,32), dK -> [32,64), dQ -> [64,96) ----
    if (warp == 4 && lane == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // K = 128 query rows, 16 per step
        umma_bf16(tmem, make_smem_desc(Pb + ks * 2048, 16384, 1024, kSwizzle128),
                  make_smem_desc(Gs + ks * 1024, 512, 512, kSwizzle64), idesc_t, ks > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        umma_bf16(tmem + 32, make_smem_desc(Db + ks * 2048, 16384, 1024, kSwizzle128),
                  make_smem_desc(Qs + ks * 1024, 512, 512, kSwizzle64), idesc_t, ks > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // K = 128 stacked keys, 16 per step: column chunk ks/4, 32 B per step inside it
        umma_bf16(tmem + 64, make_smem_desc(Db + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, kSwizzle128),
                  make_smem_desc(Ks + ks * 1024, 512, 512, kSwizzle64), idesc_q, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
    if (warp < 4) {
      mbar_wait(bar_o, it & 1);
      tc_fence_after();
      const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
      uint32_t a[32];
      // dV (row = key)
      tmem_ld32(taddr, a);
      tmem_ld_wait();
      if (tok >= 0) {
        float4* dst = reinterpret_cast<float4*>(P.dv + static_cast<int64_t>(tok) * C + h * HD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v = make_float4(__uint_as_float(a[4 * c]), __uint_as_float(a[4 * c + 1]), __uint_as_float(a[4 * c + 2]),
                                 __uint_as_float(a[4 * c + 3]));
          if (P.dv_acc) {
            const float4 o = dst[c];
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          dst[c] = v;
        }
      }
      // dK (row = key)
      tmem_ld32(taddr + 32, a);
      tmem_ld_wait();
      if (tok >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(P.dqk + static_cast<int64_t>(tok) * 2 * C + C + h * HD);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(pack_bf16(__uint_as_float(a[8 * c]), __uint_as_float(a[8 * c + 1])),
                              pack_bf16(__uint_as_float(a[8 * c + 2]), __uint_as_float(a[8 * c + 3])),
                              pack_bf16(__uint_as_float(a[8 * c + 4]), __uint_as_float(a[8 * c + 5])),
                              pack_bf16(__uint_as_float(a[8 * c + 6]), __uint_as_float(a[8 * c + 7])));
      } else if (tok == -1) {  // zero-padded key: k == bias, so its gradient goes to the k half of qk.bias
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(P.d_qk_bias + C + h * HD + j, __uint_as_float(a[j]));
      }
      // dQ (row = query); d(xW+b) = dq * scale because q was stored pre-scaled
      tmem_ld32(taddr + 64, a);
      tmem_ld_wait();
      if (tok >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(P.dqk + static_cast<int64_t>(tok) * 2 * C + h * HD);
        const float sc = P.scale;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(pack_bf16(sc * __uint_as_float(a[8 * c]), sc * __uint_as_float(a[8 * c + 1])),
                              pack_bf16(sc * __uint_as_float(a[8 * c + 2]), sc * __uint_as_float(a[8 * c + 3])),
                              pack_bf16(sc * __uint_as_float(a[8 * c + 4]), sc * __uint_as_float(a[8 * c + 5])),
                              pack_bf16(sc * __uint_as_float(a[8 * c + 6]), sc * __uint_as_float(a[8 * c + 7])));
      }
    }
  }

  // flush the relative-position-bias gradient: dTable[idx(i,j), h] += sum_windows dS[i][j]
  __syncthreads();
  float* dt = tbl;  // reuse as the CTA-level accumulator
  for (int i = threadIdx.x; i < 176; i += kAttnThreads) dt[i] = 0.f;
  __syncthreads();
  if (warp < 4 && my_pos < kNTok) {
    const int bi = rpb_base(my_pos);
#pragma unroll
    for (int j = 0; j < kNTok; ++j) atomicAdd(dt + bi - rpb_col(j), dtab[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 169; i += kAttnThreads) atomicAdd(P.d_table + i * P.nH + h, dt[i]);

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace

int launch_attn_fwd_pipe(const AttnParams& P, const crf_block_desc& d, cudaStream_t st);
int launch_attn_bwd_pipe(const AttnParams& P, const crf_block_desc& d, cudaStream_t st);
int launch_attn_fwd_async(const AttnParams& P, const crf_block_desc& d, cudaStream_t st);
int launch_attn_bwd_async(const AttnParams& P, const crf_block_desc& d, cudaStream_t st);
static bool use_legacy_attn() {
  static const bool v = getenv("CRF_ATTN_LEGACY") != nullptr;  // development switch: single-buffer kernels
  return v;
}
// development switch for A/B measurements: CRF_ATTN_IMPL=pipe selects the second-generation kernels (crf_attn_pipe.cu)
static bool use_pipe_attn() {
  static const bool v = getenv("CRF_ATTN_IMPL") != nullptr && getenv("CRF_ATTN_IMPL")[0] == 'p';
  return v;
}

int launch_attn_fwd(const crf_block_desc& d, const void* qk, const void* vb, const float* qk_bias, float scale,
                    const float* table, const float* ext_mask, int ext_mask_nw, void* o, float* lse, cudaStream_t st) {
  AttnParams P{};
  if (fill_attn_params(P, d)) return 1;
  P.qk = reinterpret_cast<const __nv_bfloat16*>(qk);
  P.vb = reinterpret_cast<const __nv_bfloat16*>(vb);
  P.qk_bias = qk_bias;
  P.table = table;
  P.scale = scale;
  P.ext_mask = ext_mask_nw > 0 ? ext_mask : nullptr;
  P.ext_mask_nw = ext_mask_nw > 0 ? ext_mask_nw : 1;
  P.o = reinterpret_cast<__nv_bfloat16*>(o);
  P.lse = lse;
  P.prof = getenv("CRF_ATTN_PROF") != nullptr;
  if (!use_legacy_attn()) return use_pipe_attn() ? launch_attn_fwd_pipe(P, d, st) : launch_attn_fwd_async(P, d, st);
  const size_t smem = 40960 + 176 * 4 + 128 + 32 + 1024;
  CRF_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int gx = (num_sms(d.device) * 4 + P.nH - 1) / P.nH;
  if (gx > P.npairs) gx = P.npairs;
  if (gx < 1) gx = 1;
  const double TC = static_cast<double>(d.B) * d.H * d.W * d.C;
  KernelTimer tm(st, 4.0 * 49 * 49 * d.C * P.total_windows, 8.0 * TC, "attn_fwd_B%d_%dx%d_C%d_s%d", d.B, d.H, d.W, d.C,
                 d.shift);
  attn_fwd_kernel<<<dim3(gx, P.nH), kAttnThreads, smem, st>>>(P);
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_attn_bwd(const crf_block_desc& d, const void* qk, const void* vb, const float* qk_bias, float scale,
                    const float* table, const float* ext_mask, int ext_mask_nw, const float* lse, const void* dout,
                    void* dqk, float* dv, int dv_acc, float* d_table, float* d_qk_bias, cudaStream_t st) {
  AttnParams P{};
  if (fill_attn_params(P, d)) return 1;
  P.qk = reinterpret_cast<const __nv_bfloat16*>(qk);
  P.vb = reinterpret_cast<const __nv_bfloat16*>(vb);
  P.qk_bias = qk_bias;
  P.table = table;
  P.scale = scale;
  P.ext_mask = ext_mask_nw > 0 ? ext_mask : nullptr;
  P.ext_mask_nw = ext_mask_nw > 0 ? ext_mask_nw : 1;
  P.lse = const_cast<float*>(lse);
  P.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  P.dqk = reinterpret_cast<__nv_bfloat16*>(dqk);
  P.dv = dv;
  P.dv_acc = dv_acc;
  P.d_table = d_table;
  P.d_qk_bias = d_qk_bias;
  P.prof = getenv("CRF_ATTN_PROF") != nullptr;
  if (!use_legacy_attn()) return use_pipe_attn() ? launch_attn_bwd_pipe(P, d, st) : launch_attn_bwd_async(P, d, st);
  const size_t smem = 98304 + 176 * 4 + 128 + 32 + 1024;
  CRF_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int gx = (num_sms(d.device) * 2 + P.nH - 1) / P.nH;
  if (gx > P.npairs) gx = P.npairs;
  if (gx < 1) gx = 1;
  const double TC = static_cast<double>(d.B) * d.H * d.W * d.C;
  KernelTimer tm(st, 10.0 * 49 * 49 * d.C * P.total_windows, 16.0 * TC, "attn_bwd_B%d_%dx%d_C%d_s%d", d.B, d.H, d.W,
                 d.C, d.shift);
  attn_bwd_kernel<<<dim3(gx, P.nH), kAttnThreads, smem, st>>>(P);
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace crf
