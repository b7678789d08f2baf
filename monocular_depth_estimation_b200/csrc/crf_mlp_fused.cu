// Fused MLP half of a CRF block, forward:  y = x1 + fc2(GELU(fc1(LayerNorm2(x1))))
// (CRFBlock.forward, /root/reference/src/newcrf_layers.py:255 with Mlp.forward :21-27), ONE kernel for C = 128 / 256.
//
// A persistent CTA per SM loops over 128-token tiles.  The 4C-wide hidden activation never goes to HBM on the way
// from fc1 to fc2: it lives in TMEM (fc1 accumulator chunks), registers (bias + GELU) and shared memory (the bf16
// A-operand chunk of fc2), 64 hidden columns at a time, while the fc2 accumulator of the tile stays in TMEM.
//
//   warps 0-3    LayerNorm-2 prologue: warp per row, fp32 rows straight from HBM (8 rows in flight per warp), row
//                statistics by warp shuffles, the normalised row written as bf16 into the swizzled UMMA A-operand
//                tile (double-buffered across tiles); training: the same smem tile is TMA-stored as `xn2`, (mean, rstd)
//                go to `stats` (both are what the backward kernels read today)
//   warps 4-11   two GELU groups (thread = accumulator row): fc1 chunk j from TMEM, + bias, exact-erf GELU, bf16 into
//                the A-operand chunk of fc2; training: pre-activation and activation chunks leave by TMA store
//   warps 12-15  final epilogue: fc2 accumulator + bias + residual x1 (TMA-loaded slab) -> y, 32-column fp32 slabs
//   warp 16      TMA producer: streams W1 / W2 chunks through a ring of four 16 KB stages; the weights stay L2-resident
//   warp 17      MMA issuer: fc1 chunk j+1 is issued before fc2 chunk j, so the GELU warps never wait for a main loop
//
// TMEM: four fc1 chunk accumulators (4 x 64 columns) + the fc2 accumulator(s) (2 x 128 at C = 128, 1 x 256 at C = 256).
// Algorithmic work per token: 16 C^2 flops; HBM bytes: 8 C (x1 in, y out) + training saves 2 C + 16 C (xn2, pre, act).
#include "crf_gemm_epi.cuh"

namespace crf {

namespace {

constexpr int TM = 128;  // tokens per tile (UMMA M)
constexpr int HC = 64;   // hidden columns per chunk (one 128-byte swizzle atom of bf16)
constexpr int kThreads = 576;
constexpr int kPreBufs = 4;
constexpr int kSlab = TM * 128;  // 16 KB: 128 rows x 128 bytes

template <int C>
struct MlpPlan {
  static constexpr int NCH = C / 64;                 // 64-column atoms per normalised row
  static constexpr int NJ = 4 * C / HC;              // hidden chunks per tile
  static constexpr int kXnTile = NCH * kSlab;        // A operand of fc1: NCH atoms of 128 rows x 128 B
  static constexpr int kXnBufs = C <= 128 ? 2 : 1;
  static constexpr int NSUB = C / 128;               // ring stages per weight chunk
  static constexpr int kStage = HC * 128 * 2;        // 16 KB: W1[chunk rows, 128 k-columns] or W2[128 rows, chunk columns]
  static constexpr int kStages = 4;
  static constexpr int kYBufs = C <= 128 ? 2 : 1;
  static constexpr int kXnOff = 0;
  static constexpr int kRingOff = kXnOff + kXnBufs * kXnTile;
  static constexpr int kActOff = kRingOff + kStages * kStage;   // 2 x 16 KB: A operand chunks of fc2
  static constexpr int kPreOff = kActOff + 2 * kSlab;           // 2 x 16 KB: staging of the pre-activation store
  static constexpr int kFinOff = kPreOff + 2 * kSlab;           // aux slab + out slab of the final epilogue
  static constexpr int kBarOff = kFinOff + 2 * kSlab;
  // barriers: xn_full/empty[2], w_full/empty[kStages], pre_full/empty[4], act_full/empty[2], y_full/empty[2], aux
  static constexpr int kNumBars = 4 + 2 * kStages + 2 * kPreBufs + 4 + 4 + 1;
  static constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024;
  static_assert(kSmemBytes <= 232448, "shared-memory plan exceeds 227 KB");
  static_assert(kPreBufs * HC + kYBufs * C <= 512, "TMEM plan exceeds 512 columns");
};

struct MlpArgs {
  const float* x1;     // (T, C) fp32 residual stream
  const float* gamma;  // LayerNorm-2 weight / bias
  const float* beta;
  const float* b1;     // (4C)
  const float* b2;     // (C)
  float* stats;        // (T, 2) (mean, rstd) or nullptr
  float eps;
  int T;
  int training;        // 1: xn2 / pre / act are stored for the backward pass
};

template <int C>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                     const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ CUtensorMap tmXn, const __grid_constant__ CUtensorMap tmPre,
                     const __grid_constant__ CUtensorMap tmAct, const MlpArgs a) {
  using PL = MlpPlan<C>;
  constexpr int NCH = PL::NCH, NJ = PL::NJ, NSUB = PL::NSUB, kStages = PL::kStages, kXnBufs = PL::kXnBufs, kYBufs = PL::kYBufs;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + PL::kBarOff;
  int bi = 0;
  const uint32_t xn_full = bar0 + 8u * bi; bi += 2;
  const uint32_t xn_empty = bar0 + 8u * bi; bi += 2;
  const uint32_t w_full = bar0 + 8u * bi; bi += kStages;
  const uint32_t w_empty = bar0 + 8u * bi; bi += kStages;
  const uint32_t pre_full = bar0 + 8u * bi; bi += kPreBufs;
  const uint32_t pre_empty = bar0 + 8u * bi; bi += kPreBufs;
  const uint32_t act_full = bar0 + 8u * bi; bi += 2;
  const uint32_t act_empty = bar0 + 8u * bi; bi += 2;
  const uint32_t y_full = bar0 + 8u * bi; bi += 2;
  const uint32_t y_empty = bar0 + 8u * bi; bi += 2;
  const uint32_t aux_bar = bar0 + 8u * bi; bi += 1;
  const uint32_t tmem_ptr_addr = bar0 + 8u * PL::kNumBars;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + PL::kBarOff + 8 * PL::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = (a.T + TM - 1) / TM;
  const int my_tiles =
      (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmY);
    for (int i = 0; i < 2; ++i) {
      mbar_init(xn_full + 8u * i, 1);
      mbar_init(xn_empty + 8u * i, 1);
      mbar_init(act_full + 8u * i, 1);
      mbar_init(act_empty + 8u * i, 1);
      mbar_init(y_full + 8u * i, 1);
      mbar_init(y_empty + 8u * i, 128);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(w_full + 8u * i, 1);
      mbar_init(w_empty + 8u * i, 1);
    }
    for (int i = 0; i < kPreBufs; ++i) {
      mbar_init(pre_full + 8u * i, 1);
      mbar_init(pre_empty + 8u * i, 128);
    }
    mbar_init(aux_bar, 1);
    fence_mbar_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t tmem_pre = tmem_base, tmem_y = tmem_base + kPreBufs * HC;

  if (warp == 16) {
    // ===== TMA producer: W1 chunk 0, then (W1 chunk s, W2 chunk s-1) ... in the order the MMA warp consumes them =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        for (int s = 0; s <= NJ; ++s) {
          if (s < NJ) {
            for (int h = 0; h < NSUB; ++h, ++it) {  // W1[s*HC .. +HC, 128h .. 128h+128): two 64-column atoms
              const uint32_t st = it % kStages;
              if (it >= static_cast<uint32_t>(kStages)) mbar_wait(w_empty + 8u * st, ((it / kStages) - 1) & 1);
              const uint32_t dst = base + PL::kRingOff + st * PL::kStage;
              mbar_expect_tx(w_full + 8u * st, PL::kStage);
              tma_load_2d(dst, &tmW1, w_full + 8u * st, 128 * h, s * HC);
              tma_load_2d(dst + HC * 128, &tmW1, w_full + 8u * st, 128 * h + 64, s * HC);
            }
          }
          if (s >= 1) {
            for (int h = 0; h < NSUB; ++h, ++it) {  // W2[128h .. 128h+128, (s-1)*HC .. +HC): one atom column
              const uint32_t st = it % kStages;
              if (it >= static_cast<uint32_t>(kStages)) mbar_wait(w_empty + 8u * st, ((it / kStages) - 1) & 1);
              const uint32_t dst = base + PL::kRingOff + st * PL::kStage;
              mbar_expect_tx(w_full + 8u * st, PL::kStage);
              tma_load_2d(dst, &tmW2, w_full + 8u * st, (s - 1) * HC, 128 * h);
            }
          }
        }
      }
    }
  } else if (warp == 17) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(1u, 0u, 0u, TM, HC);
      const uint32_t idesc2 = make_idesc(1u, 0u, 0u, TM, 128);
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int xb = i % kXnBufs, yb = i % kYBufs;
        mbar_wait(xn_full + 8u * xb, (i / kXnBufs) & 1);
        tc_fence_after();
        const uint32_t xn_s = base + PL::kXnOff + xb * PL::kXnTile;
        for (int s = 0; s <= NJ; ++s) {
          if (s < NJ) {  // fc1 chunk s: pre[128 x HC] = xn[128 x C] * W1[s*HC.., :]^T
            const uint32_t g = static_cast<uint32_t>(i) * NJ + s, pb = g % kPreBufs;
            if (g >= static_cast<uint32_t>(kPreBufs)) {
              mbar_wait(pre_empty + 8u * pb, ((g / kPreBufs) - 1) & 1);
              tc_fence_after();
            }
            const SmemDescBase ad = make_smem_desc_base(xn_s, 16, 1024, kSwizzle128);
            for (int h = 0; h < NSUB; ++h, ++it) {
              const uint32_t st = it % kStages;
              mbar_wait(w_full + 8u * st, (it / kStages) & 1);
              tc_fence_after();
              const SmemDescBase bd = make_smem_desc_base(base + PL::kRingOff + st * PL::kStage, 16, 1024, kSwizzle128);
#pragma unroll
              for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_bf16(tmem_pre + pb * HC, smem_desc_at(ad, (2 * h + k) * kSlab + ks * 32),
                            smem_desc_at(bd, k * (HC * 128) + ks * 32), idesc1, (h > 0 || k > 0 || ks > 0) ? 1u : 0u);
              umma_commit(w_empty + 8u * st);
            }
            umma_commit(pre_full + 8u * pb);
            if (s == NJ - 1) umma_commit(xn_empty + 8u * xb);
          }
          if (s >= 1) {  // fc2 chunk j: y[128 x C] += act_j[128 x HC] * W2[:, j*HC..]^T
            const int j = s - 1;
            const uint32_t g = static_cast<uint32_t>(i) * NJ + j, ab = g & 1u;
            mbar_wait(act_full + 8u * ab, (g >> 1) & 1);
            tc_fence_after();
            if (j == 0 && i >= kYBufs) {
              mbar_wait(y_empty + 8u * yb, ((i / kYBufs) - 1) & 1);
              tc_fence_after();
            }
            const SmemDescBase ad = make_smem_desc_base(base + PL::kActOff + ab * kSlab, 16, 1024, kSwizzle128);
            for (int h = 0; h < NSUB; ++h, ++it) {  // output columns [128h, 128h + 128)
              const uint32_t st = it % kStages;
              mbar_wait(w_full + 8u * st, (it / kStages) & 1);
              tc_fence_after();
              const SmemDescBase bd = make_smem_desc_base(base + PL::kRingOff + st * PL::kStage, 16, 1024, kSwizzle128);
#pragma unroll
              for (int ks = 0; ks < HC / 16; ++ks)
                umma_bf16(tmem_y + yb * C + h * 128, smem_desc_at(ad, ks * 32), smem_desc_at(bd, ks * 32), idesc2,
                          (j > 0 || ks > 0) ? 1u : 0u);
              umma_commit(w_empty + 8u * st);
            }
            umma_commit(act_empty + 8u * ab);
            if (j == NJ - 1) umma_commit(y_full + 8u * yb);
          }
        }
      }
    }
  } else if (warp < 4) {
    // ===== LayerNorm-2 prologue: rows 32*warp .. 32*warp+31 of every tile, 8 rows in flight =====
    const int tid = threadIdx.x;  // 0..127
    float2 gam[NCH], bet[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      gam[k] = __ldg(reinterpret_cast<const float2*>(a.gamma + 64 * k + 2 * lane));
      bet[k] = __ldg(reinterpret_cast<const float2*>(a.beta + 64 * k + 2 * lane));
    }
    for (int i = 0; i < my_tiles; ++i) {
      const int t0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TM;
      const int xb = i % kXnBufs;
      if (tid == 0) {
        if (i >= kXnBufs) mbar_wait(xn_empty + 8u * xb, ((i / kXnBufs) - 1) & 1);  // fc1 of tile i-kXnBufs has read it
        if (a.training) bulk_wait_read<0>();                                        // and so has its xn2 store
      }
      named_bar_sync(1, 128);
      uint8_t* xn_g = gen + PL::kXnOff + xb * PL::kXnTile;
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 8) {
        float2 v[8][NCH];
        float s[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int t = t0 + 32 * warp + r0 + q;
          const int tc = t < a.T ? t : a.T - 1;  // tail rows re-read the last row (zeroed below, never stored)
          const float* row = a.x1 + static_cast<size_t>(tc) * C;
          s[q] = 0.f;
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            v[q][k] = __ldg(reinterpret_cast<const float2*>(row + 64 * k + 2 * lane));
            s[q] += v[q][k].x + v[q][k].y;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = 32 * warp + r0 + q, t = t0 + r;
          const float mean = warp_sum(s[q]) * (1.0f / C);
          float d2 = 0.f;
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            const float d0 = v[q][k].x - mean, d1 = v[q][k].y - mean;
            d2 += d0 * d0 + d1 * d1;
          }
          const float rstd = rsqrtf(warp_sum(d2) * (1.0f / C) + a.eps);
          const bool live = t < a.T;
          if (live && lane == 0 && a.stats != nullptr)
            *reinterpret_cast<float2*>(a.stats + 2 * static_cast<size_t>(t)) = make_float2(mean, rstd);
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            const float a0 = (v[q][k].x - mean) * rstd * gam[k].x + bet[k].x;
            const float a1 = (v[q][k].y - mean) * rstd * gam[k].y + bet[k].y;
            // atom k, row r, bf16 columns 2*lane, 2*lane+1: 16-byte chunk lane/4, 4-byte slot lane%4
            *reinterpret_cast<uint32_t*>(xn_g + k * kSlab + sw128_offset(r, lane >> 2) + ((lane & 3) << 2)) =
                live ? pack_bf16(a0, a1) : 0u;
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (tid == 0) {
        mbar_arrive(xn_full + 8u * xb);
        if (a.training) {
#pragma unroll
          for (int k = 0; k < NCH; ++k) tma_store_2d(&tmXn, base + PL::kXnOff + xb * PL::kXnTile + k * kSlab, 64 * k, t0);
          bulk_commit();
        }
      }
    }
    if (tid == 0) bulk_wait_read<0>();
  } else if (warp < 12) {
    // ===== GELU groups: group gi takes the chunks with (global) parity gi =====
    const int gi = (warp - 4) >> 2;
    const int r = threadIdx.x & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint8_t* act_g = gen + PL::kActOff + gi * kSlab;
    uint8_t* pre_g = gen + PL::kPreOff + gi * kSlab;
    const uint32_t act_s = base + PL::kActOff + gi * kSlab, pre_s = base + PL::kPreOff + gi * kSlab;
    for (int i = 0; i < my_tiles; ++i) {
      const int t0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TM;
#pragma unroll 1
      for (int j = gi; j < NJ; j += 2) {
        const uint32_t g = static_cast<uint32_t>(i) * NJ + j, pb = g % kPreBufs;
        mbar_wait(pre_full + 8u * pb, (g / kPreBufs) & 1);
        tc_fence_after();
        if (r == 0 && a.training) bulk_wait_read<0>();  // this group's previous pre / act stores have read their tiles
        if (g >= 2) mbar_wait(act_empty + 8u * gi, ((g >> 1) - 1) & 1);  // fc2 of chunk g-2 has read the act tile
        named_bar_sync(2 + gi, 128);
#pragma unroll
        for (int half = 0; half < HC / 32; ++half) {
          uint32_t acc[32];
          tmem_ld32(tmem_pre + pb * HC + half * 32 + lane_base, acc);
          tmem_ld_wait();
          if (half == HC / 32 - 1) {  // accumulator chunk fully read: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            mbar_arrive(pre_empty + 8u * pb);
          }
          float v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(acc[q]);
          add_bias32(v, a.b1, j * HC + half * 32);
          if (a.training) store_bf16_32(pre_g, r, half * 4, v);
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = gelu_erf(v[q]);
          store_bf16_32(act_g, r, half * 4, v);
        }
        fence_proxy_async_smem();
        named_bar_sync(2 + gi, 128);
        if (r == 0) {
          mbar_arrive(act_full + 8u * gi);
          if (a.training) {
            tma_store_2d(&tmPre, pre_s, j * HC, t0);
            tma_store_2d(&tmAct, act_s, j * HC, t0);
            bulk_commit();
          }
        }
      }
    }
    if (r == 0) bulk_wait_read<0>();
  } else {
    // ===== final epilogue: y = fc2 accumulator + b2 + x1, 32-column fp32 slabs =====
    const int r = threadIdx.x & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t aux_s = base + PL::kFinOff, out_s = aux_s + kSlab;
    uint8_t* auxb = gen + PL::kFinOff;
    uint8_t* outb = auxb + kSlab;
    EpiParams ep{a.b2, 1.f, 0, 0, 1, nullptr, 0};
    int aux_cnt = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int t0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TM;
      const int yb = i % kYBufs;
      if (r == 0) {  // the aux slab was consumed before the last barrier of the previous tile
        mbar_expect_tx(aux_bar, kSlab);
        tma_load_2d(aux_s, &tmX1, aux_bar, 0, t0);
      }
      mbar_wait(y_full + 8u * yb, (i / kYBufs) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < C / 32; ++s) {
        if (r == 0) bulk_wait_read<0>();  // the previous store has read the out slab
        named_bar_sync(4, 128);
        mbar_wait(aux_bar, aux_cnt & 1);
        ++aux_cnt;
        uint32_t acc[32];
        tmem_ld32(tmem_y + yb * C + s * 32 + lane_base, acc);
        tmem_ld_wait();
        if (s == C / 32 - 1) {
          tc_fence_before();
          mbar_arrive(y_empty + 8u * yb);
        }
        epi_group32<CRF_EPI_BIAS_RES_F32>(acc, ep, s * 32, r, 0, outb, auxb);
        fence_proxy_async_smem();
        named_bar_sync(4, 128);
        if (r == 0) {
          tma_store_2d(&tmY, out_s, s * 32, t0);
          bulk_commit();
          if (s + 1 < C / 32) {
            mbar_expect_tx(aux_bar, kSlab);
            tma_load_2d(aux_s, &tmX1, aux_bar, (s + 1) * 32, t0);
          }
        }
      }
    }
    if (r == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int C>
int launch_mlp_c(const crf_mlp_args& m, cudaStream_t st) {
  using PL = MlpPlan<C>;
  CUtensorMap tmW1, tmW2, tmX1, tmY, tmXn, tmPre, tmAct;
  if (make_tmap_bf16(&tmW1, m.w1_bf16, 4 * C, C, HC)) return 1;
  if (make_tmap_bf16(&tmW2, m.w2_bf16, C, 4 * C, 128)) return 1;
  if (make_tmap_f32(&tmX1, m.x1, m.T, C, TM)) return 1;
  if (make_tmap_f32(&tmY, m.y, m.T, C, TM)) return 1;
  if (m.training) {
    if (make_tmap_bf16(&tmXn, m.xn2, m.T, C, TM)) return 1;
    if (make_tmap_bf16(&tmPre, m.pre, m.T, 4 * C, TM)) return 1;
    if (make_tmap_bf16(&tmAct, m.act, m.T, 4 * C, TM)) return 1;
  } else {
    tmXn = tmW1; tmPre = tmW1; tmAct = tmW1;
  }
  auto kern = mlp_fused_fwd_kernel<C>;
  CRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PL::kSmemBytes));
  const int tiles = (m.T + TM - 1) / TM;
  int grid = num_sms(m.device);
  if (grid > tiles) grid = tiles;
  MlpArgs a{m.x1, m.norm_w, m.norm_b, m.b1, m.b2, m.training ? m.stats : nullptr, m.eps, m.T, m.training ? 1 : 0};
  const double tc = static_cast<double>(m.T) * C;
  KernelTimer tm(st, 16.0 * tc * C, tc * (8.0 + (m.training ? 18.0 : 0.0)) + 16.0 * C * C, "mlp_fused_fwd_T%d_C%d%s",
                 m.T, C, m.training ? "" : "_infer");
  kern<<<grid, kThreads, PL::kSmemBytes, st>>>(tmW1, tmW2, tmX1, tmY, tmXn, tmPre, tmAct, a);
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace

bool mlp_fused_supported(int C) { return C == 128 || C == 256; }

int launch_mlp_fused_fwd(const crf_mlp_args& m, cudaStream_t st) {
  CRF_CHECK(m.x1 && m.y && m.w1_bf16 && m.w2_bf16 && m.b1 && m.b2 && m.norm_w && m.norm_b, "crf_mlp_fwd: null pointer");
  CRF_CHECK(m.T > 0, "crf_mlp_fwd: empty input");
  CRF_CHECK(!m.training || (m.xn2 && m.stats && m.pre && m.act), "crf_mlp_fwd: training needs xn2, stats, pre, act");
  switch (m.C) {
    case 128: return launch_mlp_c<128>(m, st);
    case 256: return launch_mlp_c<256>(m, st);
    default: return set_error("crf_mlp_fwd: C=%d is not supported by the fused kernel (128, 256)", m.C);
  }
}

}  // namespace crf
