// TMA-fed tcgen05 GEMM used for every dense projection of the CRF block, forward and backward:
//   D[M,N] = sum_k A(m,k) * B(n,k),  bf16 operands, fp32 accumulation in TMEM.
//
//   fprop  (qk, proj, fc1, fc2)    : A = activations (tokens, Cin)  K-major, B = weight (Cout, Cin) K-major
//   dgrad  (dX = dY * W)           : A = dY (tokens, Cout)          K-major, B = weight (Cout, Cin) MN-major
//   wgrad  (dW = dY^T * X)         : A = dY (tokens, Cout) MN-major, B = X (tokens, Cin) MN-major, split-K over tokens
//
// Both orientations read the SAME smem tile format: rows of 64 bf16 (128 B) with the 128-byte swizzle, as
// written by a TMA box (64 cols x R rows).  A K-major operand uses one box of R = tile rows; an MN-major
// operand uses (tile MN extent / 64) boxes of 64 K-rows each, 8 KB apart (the descriptor's leading byte offset).
//
// CTA = 6 warps: warps 0-3 epilogue (thread = accumulator row = TMEM lane), warp 4 TMA producer, warp 5 MMA
// issuer + TMEM allocator.  One 128 x BN output tile per CTA, K streamed in 64-element chunks through a
// multi-stage mbarrier ring.  Two CTAs are co-resident per SM (<= 256 TMEM columns, <= ~100 KB smem each) so
// one CTA's epilogue overlaps the other's main loop.
#include "crf_host.h"
#include "crf_ptx.cuh"

namespace crf {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kATileBytes = BM * 128;  // 16 KB

struct EpiParams {
  void* out0;
  void* out1;
  const float* bias;
  const void* aux1;
  int64_t ld;
  float scale;
  int scale_cols;
};

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, int row, int n, const uint32_t (&r)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const int64_t off = static_cast<int64_t>(row) * ep.ld + n;

  if constexpr (EPI == CRF_EPI_ATOMIC_F32) {
    float* o = reinterpret_cast<float*>(ep.out0) + off;
#pragma unroll
    for (int j = 0; j < 32; ++j) atomicAdd(o + j, v[j]);
    return;
  }

  if (ep.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }

  if constexpr (EPI == CRF_EPI_STORE_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out0) + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else if constexpr (EPI == CRF_EPI_STORE_BF16) {
    if (n < ep.scale_cols) {  // scale_cols is a multiple of 32, so the whole chunk is on one side
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= ep.scale;
    }
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out0) + off);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                        pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
  } else if constexpr (EPI == CRF_EPI_BIAS_RES_F32) {
    const float4* res = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.aux1) + off);
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out0) + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 q = __ldg(res + j);
      o[j] = make_float4(v[4 * j] + q.x, v[4 * j + 1] + q.y, v[4 * j + 2] + q.z, v[4 * j + 3] + q.w);
    }
  } else if constexpr (EPI == CRF_EPI_BIAS_GELU) {
    if (ep.out0 != nullptr) {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out0) + off);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    uint4* o1 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out1) + off);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o1[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
  } else if constexpr (EPI == CRF_EPI_MUL_DGELU) {
    const uint4* pre = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(ep.aux1) + off);
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out0) + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 p = __ldg(pre + j);
      const float g0 = v[8 * j + 0] * dgelu_erf(bf16_lo(p.x)), g1 = v[8 * j + 1] * dgelu_erf(bf16_hi(p.x));
      const float g2 = v[8 * j + 2] * dgelu_erf(bf16_lo(p.y)), g3 = v[8 * j + 3] * dgelu_erf(bf16_hi(p.y));
      const float g4 = v[8 * j + 4] * dgelu_erf(bf16_lo(p.z)), g5 = v[8 * j + 5] * dgelu_erf(bf16_hi(p.z));
      const float g6 = v[8 * j + 6] * dgelu_erf(bf16_lo(p.w)), g7 = v[8 * j + 7] * dgelu_erf(bf16_hi(p.w));
      o[j] = make_uint4(pack_bf16(g0, g1), pack_bf16(g2, g3), pack_bf16(g4, g5), pack_bf16(g6, g7));
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
            int total_chunks, int chunks_per_split, int stages, int a_major, int b_major, EpiParams ep) {
  constexpr int kBTileBytes = BN * 128;
  constexpr int kStageBytes = kATileBytes + kBTileBytes;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int kc_begin = blockIdx.z * chunks_per_split;
  const int kc_end = min(total_chunks, kc_begin + chunks_per_split);
  const int nk = kc_end - kc_begin;
  if (nk <= 0) return;  // uniform for the whole CTA

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + stages * kStageBytes;  // full[stages], empty[stages], tmem_full, tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * stages);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * stages + 1);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + stages * kStageBytes + 8 * (2 * stages + 1));

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_ptr_addr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % stages;
        if (i >= stages) mbar_wait(empty_bar(s), ((i / stages) - 1) & 1);
        const uint32_t a_dst = smem_base + s * kStageBytes;
        const uint32_t b_dst = a_dst + kATileBytes;
        mbar_expect_tx(full_bar(s), kStageBytes);
        const int k0 = (kc_begin + i) * BK;
        if (a_major == 0) {
          tma_load_2d(a_dst, &tmA, full_bar(s), k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, full_bar(s), m0 + 64 * j, k0);
        }
        if (b_major == 0) {
          tma_load_2d(b_dst, &tmB, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, full_bar(s), n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(1u, static_cast<uint32_t>(a_major), static_cast<uint32_t>(b_major), BM, BN);
      for (int i = 0; i < nk; ++i) {
        const int s = i % stages;
        mbar_wait(full_bar(s), (i / stages) & 1);
        tc_fence_after();
        const uint32_t a_src = smem_base + s * kStageBytes;
        const uint32_t b_src = a_src + kATileBytes;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t ad = (a_major == 0) ? make_smem_desc(a_src + ks * 32, 16, 1024, kSwizzle128)
                                             : make_smem_desc(a_src + ks * 2048, 8192, 1024, kSwizzle128);
          const uint64_t bd = (b_major == 0) ? make_smem_desc(b_src + ks * 32, 16, 1024, kSwizzle128)
                                             : make_smem_desc(b_src + ks * 2048, 8192, 1024, kSwizzle128);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: thread = accumulator row =====
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m0 + threadIdx.x;
    const bool row_ok = row < M;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c, r);
      tmem_ld_wait();
      if (row_ok) epilogue_chunk<EPI>(ep, row, n0 + c, r);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int EPI>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const crf_gemm_args& a, int total_chunks,
               int chunks_per_split, int splits, cudaStream_t st) {
  constexpr int kStageBytes = kATileBytes + BN * 128;
  // 2 co-resident CTAs/SM when K is short; one deep ring when K is long.
  int stages = (total_chunks / splits >= 8) ? (BN == 256 ? 4 : 6) : (BN == 256 ? 2 : 3);
  if (stages > chunks_per_split) stages = chunks_per_split < 2 ? 2 : chunks_per_split;
  const size_t smem = static_cast<size_t>(stages) * kStageBytes + 1024 + 8 * (2 * stages + 2);
  auto kern = gemm_kernel<BN, EPI>;
  CRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  EpiParams ep{a.out0, a.out1, a.bias, a.aux1, a.ld_out, a.scale, a.scale_cols};
  dim3 grid((a.M + BM - 1) / BM, a.N / BN, splits);
  const double mn = static_cast<double>(a.M) * a.N;
  const double out_bytes = EPI == CRF_EPI_STORE_BF16 ? 2 * mn
                           : EPI == CRF_EPI_BIAS_RES_F32 ? 8 * mn
                           : EPI == CRF_EPI_BIAS_GELU ? (a.out0 != nullptr ? 4 * mn : 2 * mn)
                           : 4 * mn;
  KernelTimer tm(st, 2.0 * mn * a.K, 2.0 * (static_cast<double>(a.M) + a.N) * a.K + out_bytes,
                 "gemm_%s_epi%d_M%d_N%d_K%d", a.a_major ? "wgrad" : (a.b_major ? "dgrad" : "fprop"), EPI, a.M, a.N, a.K);
  kern<<<grid, kThreads, smem, st>>>(tmA, tmB, a.M, a.N, total_chunks, chunks_per_split, stages, a.a_major,
                                     a.b_major, ep);
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

template <int BN>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const crf_gemm_args& a, int tc, int cps, int splits,
              cudaStream_t st) {
  switch (a.epilogue) {
    case CRF_EPI_STORE_F32: return launch_one<BN, CRF_EPI_STORE_F32>(tmA, tmB, a, tc, cps, splits, st);
    case CRF_EPI_STORE_BF16: return launch_one<BN, CRF_EPI_STORE_BF16>(tmA, tmB, a, tc, cps, splits, st);
    case CRF_EPI_BIAS_RES_F32: return launch_one<BN, CRF_EPI_BIAS_RES_F32>(tmA, tmB, a, tc, cps, splits, st);
    case CRF_EPI_BIAS_GELU: return launch_one<BN, CRF_EPI_BIAS_GELU>(tmA, tmB, a, tc, cps, splits, st);
    case CRF_EPI_MUL_DGELU: return launch_one<BN, CRF_EPI_MUL_DGELU>(tmA, tmB, a, tc, cps, splits, st);
    case CRF_EPI_ATOMIC_F32: return launch_one<BN, CRF_EPI_ATOMIC_F32>(tmA, tmB, a, tc, cps, splits, st);
    default: return set_error("crf_gemm: unknown epilogue %d", a.epilogue);
  }
}

}  // namespace

int launch_gemm(const crf_gemm_args& a, cudaStream_t st) {
  CRF_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "crf_gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  CRF_CHECK(a.N % 64 == 0, "crf_gemm: N=%d must be a multiple of 64", a.N);
  CRF_CHECK(a.a_major == 0 || a.a_major == 1, "crf_gemm: bad a_major");
  CRF_CHECK(a.b_major == 0 || a.b_major == 1, "crf_gemm: bad b_major");
  CRF_CHECK(a.a_major == 1 || a.K % 8 == 0, "crf_gemm: K-major A needs K %% 8 == 0 (K=%d)", a.K);
  CRF_CHECK(a.a_major == 0 || a.M % 8 == 0, "crf_gemm: MN-major A needs M %% 8 == 0 (M=%d)", a.M);
  CRF_CHECK(a.split_k >= 1, "crf_gemm: split_k must be >= 1");
  CRF_CHECK(a.split_k == 1 || a.epilogue == CRF_EPI_ATOMIC_F32, "crf_gemm: split_k needs the atomic epilogue");
  const int BN = (a.N % 256 == 0) ? 256 : (a.N % 128 == 0 ? 128 : 64);

  CUtensorMap tmA, tmB;
  if (a.a_major == 0) {
    if (make_tmap_bf16(&tmA, a.A, a.M, a.K, BM)) return 1;
  } else {
    if (make_tmap_bf16(&tmA, a.A, a.K, a.M, 64)) return 1;
  }
  if (a.b_major == 0) {
    if (make_tmap_bf16(&tmB, a.B, a.N, a.K, BN)) return 1;
  } else {
    if (make_tmap_bf16(&tmB, a.B, a.K, a.N, 64)) return 1;
  }
  const int total_chunks = (a.K + BK - 1) / BK;
  int splits = a.split_k > total_chunks ? total_chunks : a.split_k;
  int cps = (total_chunks + splits - 1) / splits;
  splits = (total_chunks + cps - 1) / cps;  // no empty split
  switch (BN) {
    case 256: return launch_bn<256>(tmA, tmB, a, total_chunks, cps, splits, st);
    case 128: return launch_bn<128>(tmA, tmB, a, total_chunks, cps, splits, st);
    default: return launch_bn<64>(tmA, tmB, a, total_chunks, cps, splits, st);
  }
}

}  // namespace crf
