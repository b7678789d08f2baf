// CTA-pair (cta_group::2) variant of the persistent token-major GEMM for the wide decoder scales (C >= 512):
// one tcgen05.mma covers a 256 x 256 output tile on the two SMs of a TPC.
//
// Why: with 128 x 128 single-CTA tiles every 2.1 MFLOP of MMA needs 32 KB of operands from L2 and 8 KB of shared-memory
// reads per K step -- the C >= 512 GEMMs sit at 350-550 TFLOP/s, L2- and smem-bandwidth-bound.
// A CTA pair shares both: each CTA loads its 128 rows of A and only HALF of the 256 N columns of B (32 KB per CTA for
// 8.4 MFLOP per pair: half the L2 traffic per flop), the tensor cores of both SMs read B from both shared memories,
// and each CTA keeps its 128 accumulator rows (x 256 columns, double-buffered: all 512 TMEM columns) in its own TMEM.
//
//   both CTAs    warp 8   TMA producer: its A rows + its B half into its own ring; the bytes are accounted on the LEADER's
//                         `full` barrier (cp.async.bulk.tensor ... cta_group::2, peer bit of the barrier address cleared)
//   leader only  warp 9   MMA issuer: tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16 per instruction); tcgen05.commit
//                         ... multicast::cluster frees the ring stage in BOTH CTAs and publishes the accumulator to both
//   both CTAs    warps 0-7  epilogue, two groups of four warps (column halves), thread = accumulator row, the same slab
//                         epilogues as the single-CTA kernel (crf_gemm_epi.cuh); TMEM hand-back = one remote mbarrier
//                         arrival per warp on the leader's `tmem_empty` barrier
#include <stdlib.h>

#include "crf_gemm_epi.cuh"
#include "crf_sched.h"

namespace crf {

namespace {

constexpr int PM = 256, PN = 256, BK = 64;   // pair tile
constexpr int kThreads = 320;                // 8 epilogue warps + TMA warp + MMA warp
constexpr int kATile = 128 * 128, kBTile = 128 * 128, kStage = kATile + kBTile;  // per CTA: 16 KB + 16 KB
constexpr int kStages = 4;
constexpr int kSlab = 128 * 128;
constexpr int kRing = kStages * kStage;                 // 128 KB
constexpr int kSlabOff = kRing;                         // 2 groups x (out slab + aux / second-output slab)
constexpr int kBarOff = kSlabOff + 4 * kSlab;           // 192 KB
constexpr int kNumBars = 2 * kStages + 2 + 2 + 2;       // full, empty, tmem_full[2], tmem_empty[2], aux[2]
constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024;
static_assert(kSmemBytes <= 232448, "shared-memory plan exceeds 227 KB");

// ---- work decomposition -------------------------------------------------------------------------------------------
// classic : pair p takes whole tiles p, p + npairs, ...  (rounds of 74 tiles: 40 / 76 / 80 / 152 / 160 tiles quantise badly)
// stream-K: the tiles x K-chunks sequence (tile-major) is cut into npairs equal contiguous ranges.  A range starts with
//           the TAIL of a tile another pair began (a "contributor" piece: the fp32 partial accumulator goes to this pair's
//           workspace slot, then a flag is released), continues with whole tiles and ends with the HEAD of a tile (an
//           "owner" piece: after its own K chunks the pair adds the partials of the pairs that hold the rest of the tile,
//           then runs the ordinary epilogue).  A contributor piece is always the FIRST piece of its pair and waits for
//           nothing, an owner piece is the LAST of its pair, so the flags are long set when they are needed and no pair
//           ever waits on a pair that waits (CTAs are dispatched in index order; all 74 pairs are resident).
// (Piece, PieceIter, sk_bound, sk_contributors: crf_sched.h -- compiled on the host by tests/test_sched_host.py)
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                 const __grid_constant__ CUtensorMap tmAux, int M, int N, int K, int b_major, EpiParams ep, int streamk,
                 int snap, float* sk_ws, uint32_t* sk_flags) {
  pdl_launch_dependents();
  using TR = EpiTraits<EPI>;
  constexpr int kSlabCols = TR::kSlabCols;
  constexpr int kNumSlabs = 128 / kSlabCols;   // per epilogue group (128 of the 256 columns)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // identical in both CTAs (same kernel, same layout)
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kBarOff;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };
  auto aux_bar = [&](int g) { return bar0 + 8u * (2 * kStages + 4 + g); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * kNumBars;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen + kBarOff + 8 * kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();             // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_tiles = N / PN;
  const int total_tiles = ((M + PM - 1) / PM) * n_tiles;
  const int nk_all = (K + BK - 1) / BK;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO0);
    if (TR::kHasOut1) tma_prefetch_desc(&tmO1);
    if (TR::kHasAux) tma_prefetch_desc(&tmAux);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);      // the leader's producer arrives (expect_tx); both CTAs' TMA bytes complete it
      mbar_init(empty_bar(s), 1);     // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);     // multicast tcgen05.commit
      mbar_init(tempty_bar(a), 16);   // 8 epilogue warps x 2 CTAs (used on the leader only)
      mbar_init(aux_bar(a), 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc_2sm(tmem_ptr_addr, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of BOTH CTAs are initialised before anything signals across the pair
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 8) {
    // ===== TMA producer (both CTAs): A rows [m0 + 128 rank, +128), B columns [n0 + 128 rank, +128) =====
    if (lane == 0) {
      int it = 0;
      PieceIter pit(nk_all, total_tiles, pair, npairs, streamk, snap);
      Piece pc;
      while (pit.next(pc)) {
        const int t = pc.t;
        const int m0 = (t / n_tiles) * PM + 128 * static_cast<int>(rank);
        const int n0 = (t % n_tiles) * PN + 128 * static_cast<int>(rank);
        for (int kc = pc.kc0; kc < pc.kc1; ++kc, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1);
          const uint32_t a_dst = base + s * kStage, b_dst = a_dst + kATile;
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * kStage);   // the bytes of both CTAs land on the leader's barrier
          tma_load_2d_2sm(a_dst, &tmA, full_bar(s), kc * BK, m0);
          if (b_major == 0) {
            tma_load_2d_2sm(b_dst, &tmB, full_bar(s), kc * BK, n0);
          } else {
            tma_load_2d_2sm(b_dst, &tmB, full_bar(s), n0, kc * BK);
            tma_load_2d_2sm(b_dst + 8192, &tmB, full_bar(s), n0 + 64, kc * BK);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (leader CTA only) =====
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(1u, 0u, static_cast<uint32_t>(b_major), PM, PN);
      int it = 0;
      PieceIter pit(nk_all, total_tiles, pair, npairs, streamk, snap);
      Piece pc;
      for (int i = 0; pit.next(pc); ++i) {
        const int a = i & 1;
        if (i >= 2) {
          mbar_wait(tempty_bar(a), ((i >> 1) - 1) & 1);   // both CTAs' epilogues have drained this accumulator
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + a * PN;
        const int kc0 = pc.kc0, kc1 = pc.kc1;
        for (int kc = kc0; kc < kc1; ++kc, ++it) {
          const int s = it % kStages;
          mbar_wait(full_bar(s), (it / kStages) & 1);
          tc_fence_after();
          const uint32_t a_src = base + s * kStage, b_src = a_src + kATile;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t ad = make_smem_desc(a_src + ks * 32, 16, 1024, kSwizzle128);
            const uint64_t bd = (b_major == 0) ? make_smem_desc(b_src + ks * 32, 16, 1024, kSwizzle128)
                                               : make_smem_desc(b_src + ks * 2048, 8192, 1024, kSwizzle128);
            umma_bf16_2sm(d_tmem, ad, bd, idesc, (kc > kc0 || ks > 0) ? 1u : 0u);
          }
          umma_commit_2sm(empty_bar(s), 3);   // frees the stage in both CTAs
        }
        umma_commit_2sm(tfull_bar(a), 3);     // accumulator complete, in both CTAs' TMEM
      }
    }
  } else {
    // ===== epilogue (both CTAs): group e drains accumulator columns [128 e, 128 e + 128) of this CTA's 128 rows =====
    const int e = warp >> 2;
    const int r = threadIdx.x & 127;
    const uint32_t out0_s = base + kSlabOff + e * 2 * kSlab, x_s = out0_s + kSlab;
    uint8_t* o0 = gen + kSlabOff + e * 2 * kSlab;
    uint8_t* xb = o0 + kSlab;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    int aux_cnt = 0;
    // stream-K partial slots: [pair][rank][group] x (32 column quads x 128 rows x float4) -- a warp's accesses are contiguous
    constexpr int kSlotFloats = 128 * 128;
    PieceIter pit(nk_all, total_tiles, pair, npairs, streamk, snap);
    Piece pc;
    for (int i = 0; pit.next(pc); ++i) {
      const int t = pc.t;
      const bool emits = pc.kc0 == 0;                  // this pair owns the tile: it runs the real epilogue
      const bool head = emits && pc.kc1 < nk_all;      // ... after adding the partials of the pairs that finish the tile
      const int a = i & 1;
      const int m0 = (t / n_tiles) * PM + 128 * static_cast<int>(rank);
      const int n0 = (t % n_tiles) * PN + 128 * e;
      const uint32_t taddr = tmem_base + a * PN + e * 128 + lane_base;
      if (TR::kHasAux && emits && r == 0) {
        mbar_expect_tx(aux_bar(e), kSlab);
        tma_load_2d(x_s, &tmAux, aux_bar(e), n0, m0);
      }
      // contributors of an owner piece: the pairs whose (first) piece covers the rest of this tile
      int n_contrib = 0;
      const float* contrib[kMaxContrib];
      if (head) {
        int q[kMaxContrib];
        n_contrib = sk_contributors(pair, npairs, total_tiles, nk_all, snap, t, pc.kc1, q);
        if (n_contrib < 0) n_contrib = 0;   // (the host only selects stream-K for shapes where this cannot happen)
        for (int k = 0; k < n_contrib; ++k) {
          const int slot = (q[k] * 2 + static_cast<int>(rank)) * 2 + e;
          if (r == 0) {
            while (ld_acquire_gpu(sk_flags + slot) == 0u) __nanosleep(64);
          }
          contrib[k] = sk_ws + static_cast<size_t>(slot) * kSlotFloats;
        }
      }
      mbar_wait(tfull_bar(a), (i >> 1) & 1);
      tc_fence_after();
      if (!emits) {
        // ---- contributor piece: bare fp32 partial accumulator -> this pair's workspace slot, then release the flag ----
        const int slot = (pair * 2 + static_cast<int>(rank)) * 2 + e;
        float4* dst = reinterpret_cast<float4*>(sk_ws + static_cast<size_t>(slot) * kSlotFloats) + r;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t acc[32];
          tmem_ld32(taddr + c * 32, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(dst + (c * 8 + j) * 128, make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                                        __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(a), 0);
        __threadfence();
        named_bar_sync(1 + e, 128);
        if (r == 0) st_release_gpu(sk_flags + slot, 1u);
        continue;
      }
#pragma unroll 1
      for (int s = 0; s < kNumSlabs; ++s) {
        const int nc = n0 + s * kSlabCols;
        if (r == 0) bulk_wait_read<0>();
        named_bar_sync(1 + e, 128);   // (also publishes thread 0's flag acquires of an owner piece to the group)
        if (TR::kHasAux) {
          mbar_wait(aux_bar(e), aux_cnt & 1);
          ++aux_cnt;
        }
#pragma unroll
        for (int half = 0; half < kSlabCols / 32; ++half) {
          uint32_t acc[32];
          tmem_ld32(taddr + s * kSlabCols + half * 32, acc);
          tmem_ld_wait();
          if (head) {
            const int c4 = (s * kSlabCols + half * 32) / 4;
            for (int k = 0; k < n_contrib; ++k) {
              const float4* src = reinterpret_cast<const float4*>(contrib[k]) + r;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 pv = __ldcg(src + (c4 + j) * 128);
                acc[4 * j + 0] = __float_as_uint(__uint_as_float(acc[4 * j + 0]) + pv.x);
                acc[4 * j + 1] = __float_as_uint(__uint_as_float(acc[4 * j + 1]) + pv.y);
                acc[4 * j + 2] = __float_as_uint(__uint_as_float(acc[4 * j + 2]) + pv.z);
                acc[4 * j + 3] = __float_as_uint(__uint_as_float(acc[4 * j + 3]) + pv.w);
              }
            }
          }
          epi_group32<EPI>(acc, ep, nc + half * 32, r, half, o0, xb);
        }
        if (s == kNumSlabs - 1) {  // this warp has read its part of the accumulator: one arrival per warp on the leader
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_bar(a), 0);
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + e, 128);
        if (r == 0) {
          if (EPI != CRF_EPI_BIAS_GELU || ep.store_out0) tma_store_2d(&tmO0, out0_s, nc, m0);
          if (TR::kHasOut1) tma_store_2d(&tmO1, x_s, nc, m0);
          bulk_commit();
          if (TR::kHasAux && s + 1 < kNumSlabs) {
            mbar_expect_tx(aux_bar(e), kSlab);
            tma_load_2d(x_s, &tmAux, aux_bar(e), nc + kSlabCols, m0);
          }
        }
      }
    }
    if (r == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer can still be read by an MMA or signalled
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int EPI>
int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO0, const CUtensorMap& tmO1,
                const CUtensorMap& tmAux, const crf_gemm_args& a, cudaStream_t st) {
  auto kern = gemm_pair_kernel<EPI>;
  CRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int tiles = ((a.M + PM - 1) / PM) * (a.N / PN);
  const int max_pairs = num_sms(a.device) / 2;
  const int nk = (a.K + BK - 1) / BK;
  int pairs = max_pairs;
  if (pairs > tiles) pairs = tiles;
  // Stream-K (see the kernel) runs when the caller hands over the partial-tile workspace (gemm_pair_streamk_bytes) and
  // a pair's share is at least a quarter of a tile's K loop (<= 5 contributors per tile).  MEASURED on B200 (round 2,
  // profiles/r02_gemm_streamk.md): bit-stable and correct, but NOT faster -- with all 74 pairs busy the time per K chunk
  // doubles (0.45 -> 0.8 us): these GEMMs are bound by a chip-wide resource on the operand path (tiles x K chunks x 64 KB
  // of TMA fills at ~5 TB/s either way), not by how the tiles pack into rounds.  So the block orchestration passes the workspace only with
  // CRF_GEMM_STREAMK=1; through the C ABI (crf_gemm) the workspace itself is the switch.
  int streamk = 0, snap = 1;
  float* sk_ws = nullptr;
  uint32_t* sk_flags = nullptr;
  if (a.workspace != nullptr && a.workspace_bytes >= gemm_pair_streamk_bytes(a.device)) {
    const double per_pair = static_cast<double>(tiles) * nk / max_pairs;
    if (per_pair * 4.0 >= nk && per_pair >= 4.0) {
      streamk = 1;
      pairs = max_pairs;
      snap = nk / 8 < 1 ? 1 : (nk / 8 > 4 ? 4 : nk / 8);
      sk_flags = reinterpret_cast<uint32_t*>(a.workspace);
      sk_ws = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + 4096);
      CRF_CUDA(cudaMemsetAsync(sk_flags, 0, 4096, st));
    }
  }
  EpiParams ep{a.bias, a.scale, a.scale_cols, 0, a.out0 != nullptr ? 1 : 0, nullptr, 0};
  const double mn = static_cast<double>(a.M) * a.N;
  const double out_bytes = EPI == CRF_EPI_STORE_BF16 ? 2 * mn
                           : EPI == CRF_EPI_BIAS_RES_F32 ? 8 * mn
                           : EPI == CRF_EPI_BIAS_GELU ? (a.out0 != nullptr ? 4 * mn : 2 * mn)
                           : 4 * mn;
  KernelTimer tm(st, 2.0 * mn * a.K, 2.0 * (static_cast<double>(a.M) + a.N) * a.K + out_bytes,
                 "gemm2%s_%s_epi%d_M%d_N%d_K%d", streamk ? "sk" : "", a.b_major ? "dgrad" : "fprop", EPI, a.M, a.N, a.K);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see launch_pdl (crf_host.h)
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  CRF_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO0, tmO1, tmAux, a.M, a.N, a.K, a.b_major, ep, streamk, snap,
                                sk_ws, sk_flags));
  note_launch();
  return 0;
}

}  // namespace

// flags (4 KB) + one 256 x 256 fp32 partial tile per SM pair
size_t gemm_pair_streamk_bytes(int device) {
  return 4096 + static_cast<size_t>(num_sms(device) / 2) * PM * PN * sizeof(float);
}

// fprop / dgrad (A K-major) with N a multiple of 256 and enough work for CTA pairs.  Returns -1 if not eligible.
int launch_gemm_pair(const crf_gemm_args& a, cudaStream_t st) {
  // verified (tests/test_gpu_stages.py, block parity) and measured on B200: 15-30 % faster than the single-CTA kernel on
  // every C >= 512 projection (profiles/r02_gemm_pair.md); CRF_GEMM_PAIR=0 falls back to the single-CTA kernel
  static const int mode = [] { const char* e = getenv("CRF_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (mode == 0) return -1;
  if (a.a_major != 0 || a.N % PN != 0 || a.epilogue == CRF_EPI_SPLITK_F32 || a.split3) return -1;
  if (a.K < 512 || a.N < 512) return -1;
  CUtensorMap tmA, tmB, tmO0, tmO1, tmAux;
  if (make_tmap_bf16(&tmA, a.A, a.M, a.K, 128)) return 1;
  if (a.b_major == 0) {
    if (make_tmap_bf16(&tmB, a.B, a.N, a.K, 128)) return 1;
  } else {
    if (make_tmap_bf16(&tmB, a.B, a.K, a.N, 64)) return 1;
  }
  tmO1 = tmA;
  tmAux = tmA;
  const int epi = a.epilogue;
  const bool out_f32 = (epi == CRF_EPI_STORE_F32 || epi == CRF_EPI_BIAS_RES_F32);
  if (epi == CRF_EPI_BIAS_GELU) {
    CRF_CHECK(a.out1 != nullptr, "crf_gemm: BIAS_GELU needs out1");
    if (a.out0 != nullptr) {
      if (make_tmap_bf16(&tmO0, a.out0, a.M, a.N, 128)) return 1;
    } else {
      tmO0 = tmA;
    }
    if (make_tmap_bf16(&tmO1, a.out1, a.M, a.N, 128)) return 1;
  } else {
    CRF_CHECK(a.out0 != nullptr, "crf_gemm: out0 is null");
    if (out_f32 ? make_tmap_f32(&tmO0, a.out0, a.M, a.N, 128) : make_tmap_bf16(&tmO0, a.out0, a.M, a.N, 128)) return 1;
  }
  if (epi == CRF_EPI_BIAS_RES_F32) {
    CRF_CHECK(a.aux1 != nullptr, "crf_gemm: BIAS_RES_F32 needs aux1 (residual)");
    if (make_tmap_f32(&tmAux, a.aux1, a.M, a.N, 128)) return 1;
  } else if (epi == CRF_EPI_MUL_DGELU) {
    CRF_CHECK(a.aux1 != nullptr, "crf_gemm: MUL_DGELU needs aux1 (pre-activation)");
    if (make_tmap_bf16(&tmAux, a.aux1, a.M, a.N, 128)) return 1;
  }
  switch (epi) {
    case CRF_EPI_STORE_F32: return launch_pair<CRF_EPI_STORE_F32>(tmA, tmB, tmO0, tmO1, tmAux, a, st);
    case CRF_EPI_STORE_BF16: return launch_pair<CRF_EPI_STORE_BF16>(tmA, tmB, tmO0, tmO1, tmAux, a, st);
    case CRF_EPI_BIAS_RES_F32: return launch_pair<CRF_EPI_BIAS_RES_F32>(tmA, tmB, tmO0, tmO1, tmAux, a, st);
    case CRF_EPI_BIAS_GELU: return launch_pair<CRF_EPI_BIAS_GELU>(tmA, tmB, tmO0, tmO1, tmAux, a, st);
    case CRF_EPI_MUL_DGELU: return launch_pair<CRF_EPI_MUL_DGELU>(tmA, tmB, tmO0, tmO1, tmAux, a, st);
    default: return set_error("crf_gemm: unknown epilogue %d", epi);
  }
}

}  // namespace crf
