// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA, TMEM alloc,
// TMEM load, commit) and the UMMA shared-memory / instruction descriptors used by every tensor-core kernel
// in this library.  Nothing here is specific to the CRF block.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace crf {

// ------------------------------------------------------------------------------------------------
// generic helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Non-blocking completion test (try_wait may suspend the thread for a system-dependent time: wrong for a thread that
// polls several barriers).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a barrier that never completes is a programming error; trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("crf: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}

// Same, for waits that are expected to last microseconds (a role that works once per tile): back off with nanosleep
// between polls so that the polling does not take issue slots from the warps doing the arithmetic.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if (clock64() - t0 > 4000000000LL) {
      printf("crf: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// smem tile -> global (tensor map), tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// smem tile += into global (tensor map): element-wise fp32 add performed by the L2, same bulk-group tracking
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// Programmatic dependent launch (launch_pdl in crf_host.h).  First statement of every kernel: let the NEXT kernel of the
// stream be scheduled as soon as all CTAs of this grid have started (its CTAs then park in their own griddepcontrol.wait
// until this grid has completed and its memory is visible), and wait for the PREVIOUS grid the same way.  Nothing of a
// kernel runs before the wait, so only launch latency and CTA scheduling overlap the previous kernel's tail; both
// instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
// The tcgen05 kernels split the two: barrier init, TMEM allocation, descriptor prefetch and the copies of PARAMETERS
// (biases, LayerNorm weights, the relative-position table: written by the optimiser, many full kernel boundaries earlier)
// into shared memory run BEFORE pdl_wait(), i.e. under the previous kernel's tail; activations, bf16 weight copies and
// everything else a neighbouring library kernel may have produced are only touched after it.
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // all but the N most recent groups have finished READING smem
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte copies for index-computed gathers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// The hardware performs one arrive on `bar` (counted against its expected arrivals: .noinc) once every cp.async this
// thread has issued so far has completed; the thread itself does not wait.
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// fire-and-forget 16-byte fp32 vector reduction into global memory (no return value, no read latency on the issuer)
__device__ __forceinline__ void red_add_f32x4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: MMA, commit
// ------------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC issue ONE tcgen05.mma over M = 256 --
// each CTA holds its 128 rows of A, its half of the N columns of B and its 128 accumulator rows in its own TMEM.
// Forms as in cute/arch/{copy_sm100_tma,mma_sm100_umma,tmem_allocator_sm100}.hpp and cutlass/arch/barrier.h.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {  // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are accounted on the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair; issued by one thread of the leader CTA
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in every CTA of `mask` once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// one arrival on the barrier at this offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM -> registers.  32x32b: warp w (w % 4 selects the 32-lane sub-partition) reads its 32 lanes,
// thread i gets lane 32*(w%4)+i, 32 consecutive fp32 columns.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same shape: thread i writes 32 consecutive fp32 columns of lane 32*(w%4)+i
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp, restated)
// ------------------------------------------------------------------------------------------------
enum : uint32_t { kSwizzleNone = 0, kSwizzle128 = 2, kSwizzle64 = 4, kSwizzle32 = 6 };

// start address, leading-dim byte offset, stride-dim byte offset (all 16-byte units), version 1, swizzle mode
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t swizzle) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(swizzle & 7) << 61;
  return d;
}

// The same descriptor split into a loop-invariant part and a byte offset: the address field is (saddr >> 4) in the low
// 14 bits, so a tile-relative offset is ONE 32-bit add on the low word (valid while base + offset stays inside shared
// memory, i.e. below 2^18).  A single thread that issues dozens of small MMAs per tile is bound by the ~15 uniform-
// datapath instructions the generic form costs per descriptor pair.
struct SmemDescBase {
  uint32_t lo, hi;
};
__device__ __forceinline__ SmemDescBase make_smem_desc_base(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                            uint32_t swizzle) {
  SmemDescBase d;
  d.lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  d.hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((swizzle & 7u) << 29);
  return d;
}
__device__ __forceinline__ uint64_t smem_desc_at(const SmemDescBase& d, uint32_t byte_off) {
  return (static_cast<uint64_t>(d.hi) << 32) | static_cast<uint64_t>(d.lo + (byte_off >> 4));
}

// kind::f16 / kind::tf32 instruction descriptor: fp32 accumulate, A/B format (1 = bf16, 2 = tf32),
// majorness (0 = K-major, 1 = MN-major), N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t M,
                                                  uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// 16-byte-chunk XOR swizzles used when threads write UMMA operand tiles by hand.
// SW128: tile rows are 128 B; chunk' = chunk ^ (row & 7).   (Swizzle<3,4,3>)
// SW64 : tile rows are  64 B; chunk' = chunk ^ ((row >> 1) & 3).   (Swizzle<2,4,3>)
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}
__device__ __forceinline__ uint32_t sw64_offset(uint32_t row, uint32_t chunk) {
  return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
}

// ------------------------------------------------------------------------------------------------
// small numeric helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Exact-erf GELU (nn.GELU default, newcrf_layers.py:171) evaluated in sigmoid form:
//   Phi(x) = (1 + erf(x / sqrt2)) / 2 = 1 / (1 + 2^(-w(x))),   w(x) = log2((1 + erf(x/sqrt2)) / erfc(x/sqrt2)),
// w is odd and w(x) / x is fitted by a quadratic in t = min(x^2, 25) (minimax over |x| <= 9, tools/fit_gelu.py):
// |gelu err| < 2.6e-5, |gelu' err| < 1.1e-4 absolute -- >= 20x below the bf16 rounding of the stored values -- for
// 9 instructions (2 MUFU) instead of the 17 of an A&S 7.1.26 erf.  The GELU epilogues are issue-bound (4C
// evaluations per token per block: ncu shows 66 % issue-slot utilisation on fc1 with the erf form), so this is
// what decides whether fc1 / dgrad-fc2 run at the HBM roofline.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_cdf(float x) {
  const float t = fminf(x * x, 25.0f);
  const float r = fmaf(fmaf(0.0010142630f, t, -0.10677572f), t, -2.3011212f);  // -w(x)/x
  return __fdividef(1.0f, 1.0f + ex2_approx(x * r));
}
__device__ __forceinline__ float gelu_erf(float x) { return x * gelu_cdf(x); }
// gelu'(x) as the exact derivative of the sigmoid form above: s + x s (1 - s) ln2 w'(x), w' = c0 + 3 c1 t + 5 c2 t^2
// (|err| < 1.1e-4 against Phi(x) + x phi(x)); two MUFU ops instead of the three of the closed form -- the
// dgrad x gelu' epilogue is MUFU-bound (16 lanes per clock and SM).
__device__ __forceinline__ float dgelu_erf(float x) {
  const float t = fminf(x * x, 25.0f);
  const float s = gelu_cdf(x);
  const float wp = fmaf(fmaf(-0.0035151686f, t, 0.22203402f), t, 1.5950157f);
  return fmaf(x * wp, fmaf(-s, s, s), s);
}

// Packed fp32x2 arithmetic (sm_100: FMUL2 / FADD2 / FFMA2, one issue slot for two IEEE operations) and the GELU above
// on a pair of values: same operations in the same order as gelu_erf, hence bit-identical results, at 5.5 instead
// of 9 issue slots per element (the GELU epilogues are issue-bound).
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t X = f2_pack(x0, x1);
  float t0, t1;
  f2_unpack(f2_mul(X, X), t0, t1);
  const uint64_t T = f2_pack(fminf(t0, 25.0f), fminf(t1, 25.0f));
  uint64_t R = f2_fma(f2_pack(0.0010142630f, 0.0010142630f), T, f2_pack(-0.10677572f, -0.10677572f));
  R = f2_fma(R, T, f2_pack(-2.3011212f, -2.3011212f));
  float e0, e1;
  f2_unpack(f2_mul(X, R), e0, e1);
  float d0, d1;
  f2_unpack(f2_add(f2_pack(ex2_approx(e0), ex2_approx(e1)), f2_pack(1.0f, 1.0f)), d0, d1);
  f2_unpack(f2_mul(X, f2_pack(rcp_approx(d0), rcp_approx(d1))), x0, x1);
}

// gelu'(x) on a pair of values with packed arithmetic: the operations of dgelu_erf in the same order (bit-identical
// results), ~11 instead of ~19 issue slots per element -- the d fc2 x gelu' epilogue is issue-bound (16 epilogue warps at
// 52 % issue utilisation cover 87 us at the 1/4 scale against an HBM time of 55 us).  Returns (gelu'(x0), gelu'(x1)) packed.
__device__ __forceinline__ uint64_t dgelu_erf2(float x0, float x1) {
  const uint64_t X = f2_pack(x0, x1);
  float t0, t1;
  f2_unpack(f2_mul(X, X), t0, t1);
  const uint64_t T = f2_pack(fminf(t0, 25.0f), fminf(t1, 25.0f));
  uint64_t R = f2_fma(f2_pack(0.0010142630f, 0.0010142630f), T, f2_pack(-0.10677572f, -0.10677572f));
  R = f2_fma(R, T, f2_pack(-2.3011212f, -2.3011212f));
  float e0, e1;
  f2_unpack(f2_mul(X, R), e0, e1);
  float d0, d1;
  f2_unpack(f2_add(f2_pack(ex2_approx(e0), ex2_approx(e1)), f2_pack(1.0f, 1.0f)), d0, d1);
  const uint64_t S = f2_pack(rcp_approx(d0), rcp_approx(d1));                       // s = gelu_cdf(x)
  uint64_t WP = f2_fma(f2_pack(-0.0035151686f, -0.0035151686f), T, f2_pack(0.22203402f, 0.22203402f));
  WP = f2_fma(WP, T, f2_pack(1.5950157f, 1.5950157f));
  const uint64_t S1 = f2_fma(S ^ 0x8000000080000000ull, S, S);                       // fmaf(-s, s, s)
  return f2_fma(f2_mul(X, WP), S1, S);                                               // fmaf(x * wp, s (1 - s), s)
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace crf
