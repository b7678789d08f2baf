// Warp-specialised, double-buffered window-attention kernels (forward and backward).
//
// Same math and tile formats as the single-buffer kernels in crf_attn.cu (see the header comment there), but the
// per-window-pair chain  gather -> MMA -> softmax math -> MMA -> store  is split over roles that run concurrently
// on TWO window pairs ("lanes" g = 0/1, each with its own smem input tiles, P/dS tiles and TMEM columns):
//
//   warps  0-3   compute group 0   thread = tile row = TMEM lane; softmax / dS math, epilogue stores
//   warps  4-7   compute group 1   (same, for the other lane)
//   warps  8-11  loaders           thread = tile row; index map + 16-byte cp.async gathers of q, k, v (, dO)
//   warp   12    MMA issuer        one thread issues every tcgen05.mma; owns the TMEM allocation
//   (warps 13-15 only fill the fourth warpgroup so setmaxnreg can re-balance registers)
//
// Hand-offs are mbarriers (counts in parentheses): full[g][slot] (128 loader arrivals) -> s_done[g]
// (tcgen05.commit) -> p_ready[g] (128 compute arrivals) -> o_done[g] (tcgen05.commit) -> t_free[g] (128 compute
// arrivals: TMEM columns drained).  Each lane has TWO input slots, so the gathers of a lane's next pair run while
// its current pair is still in the tensor / softmax stages; a slot is handed back to the loaders by in_free[g][slot]
// (a second tcgen05.commit after the pair's last MMA).  The MMA thread issues S(i), then the second-stage MMAs of
// pair i-1, so one lane's tensor work overlaps the other lane's softmax math.  The backward needs 64 KB of
// block-diagonal P / dS tiles per lane, so it runs with ONE input slot per lane (sharing one tile set between the
// lanes was measured slower: it chains every pair's tile write behind the previous pair's MMAs).
#include "crf_attn_common.cuh"

namespace crf {

int fill_attn_params(AttnParams& P, const crf_block_desc& d) {
  CRF_CHECK(d.C % d.num_heads == 0 && d.C / d.num_heads == 32,
            "attention core: head_dim must be 32 (C=%d, heads=%d)", d.C, d.num_heads);
  CRF_CHECK(d.window == 7, "attention core: window must be 7 (got %d)", d.window);
  CRF_CHECK(d.shift >= 0 && d.shift < d.window, "shift_size must in 0-window_size");
  P.gm = WindowGeom(d.H, d.W, d.window, d.shift);
  P.B = d.B;
  P.C = d.C;
  P.nH = d.num_heads;
  P.total_windows = d.B * P.gm.nW;
  P.npairs = (P.total_windows + 1) / 2;
  P.rcp_nW = 1.0f / static_cast<float>(P.gm.nW);
  P.rcp_nWw = 1.0f / static_cast<float>(P.gm.nWw);
  CRF_CHECK(P.total_windows < (1 << 22), "attention core: too many windows (%d)", P.total_windows);
  return 0;
}

namespace {

constexpr int kPipeThreads = 512;
constexpr int HD = 32;

// ---- shared-memory plan (offsets from the 1024-aligned base) ----
// backward: in[g][0] = Q,K,V,dO (4 x 8 KB); P/dS tiles (2 x 32 KB) per lane: lane 0 in the tile area, lane 1 in the
// input slots [0][1] and [1][1], which the backward leaves unused
constexpr uint32_t kBwdIn = 0;                         // 4 x 32 KB
constexpr uint32_t kBwdTiles = 131072;                 // 64 KB
constexpr uint32_t kBwdMisc = 131072 + 65536;          // 196608
// forward: in[g][slot] = Q,K,V (3 x 8 KB, padded to 32 KB) x 4; tiles[g] = P (16 KB)
constexpr uint32_t kFwdIn = 0;                         // 4 x 32 KB
constexpr uint32_t kFwdTiles = 131072;                 // 2 x 16 KB
constexpr uint32_t kFwdMisc = 131072 + 32768;          // 163840
// misc block: tbl[176] f32 | tok[4][128] i32 | wgl[4][128] i32 | rid[4][128] u8 | barriers[17] | tmem ptr
constexpr uint32_t kMiscTbl = 0;
constexpr uint32_t kMiscTok = 704;
constexpr uint32_t kMiscWg = 704 + 2048;
constexpr uint32_t kMiscRid = 704 + 4096;
constexpr uint32_t kMiscBar = 704 + 4096 + 512;        // 5312, 8-byte aligned
constexpr uint32_t kMiscTmem = kMiscBar + 17 * 8;
constexpr uint32_t kMiscBytes = kMiscTmem + 16;

// backward tile placement: lane 0 uses the dedicated tile area, lane 1 the two input slots the backward does not use
__host__ __device__ constexpr uint32_t bwd_p_tile(int g) { return g == 0 ? kBwdTiles : kBwdIn + 1 * 32768; }
__host__ __device__ constexpr uint32_t bwd_ds_tile(int g) { return g == 0 ? kBwdTiles + 32768 : kBwdIn + 3 * 32768; }

struct Bars {
  uint32_t base;
  __device__ uint32_t full(int g, int slot) const { return base + 8u * (2 * g + slot); }
  __device__ uint32_t s_done(int g) const { return base + 8u * (4 + g); }
  __device__ uint32_t p_ready(int g) const { return base + 8u * (6 + g); }
  __device__ uint32_t o_done(int g) const { return base + 8u * (8 + g); }
  __device__ uint32_t t_free(int g) const { return base + 8u * (10 + g); }
  __device__ uint32_t in_free(int g, int slot) const { return base + 8u * (12 + 2 * g + slot); }
  __device__ uint32_t tiles_free() const { return base + 8u * 16; }
};

__device__ __forceinline__ void init_bars(const Bars& b) {
  for (int g = 0; g < 2; ++g) {
    for (int slot = 0; slot < 2; ++slot) {
      mbar_init(b.full(g, slot), 128);
      mbar_init(b.in_free(g, slot), 1);
    }
    mbar_init(b.s_done(g), 1);
    mbar_init(b.p_ready(g), 128);
    mbar_init(b.o_done(g), 1);
    mbar_init(b.t_free(g), 128);
  }
  mbar_init(b.tiles_free(), 1);
  fence_mbar_init();
}

// Loader body shared by forward and backward: fills in[g] for `pair` and publishes per-row metadata.
// (pi, pj) = (pos / 7, pos % 7) are per-thread constants; only two integer divisions remain per pair.
template <bool BWD>
__device__ __forceinline__ void load_pair(const AttnParams& P, int pair, int r, int pi, int pj, int h, uint32_t in_s,
                                          uint8_t* in_g, int* tok_out, int* wg_out, uint8_t* rid_out) {
  const int half = r >> 6, pos = r & 63;
  const int wg = 2 * pair + half;
  const int C = P.C;
  int tok = -2, region = 0;
  if (pos < kNTok && wg < P.total_windows) {
    const int b = wg / P.gm.nW;
    const int win = wg - b * P.gm.nW;
    const int wh = win / P.gm.nWw, ww = win - wh * P.gm.nWw;
    const int hs = wh * 7 + pi, wsx = ww * 7 + pj;  // position on the shifted, padded grid
    int hh = hs + P.gm.shift, wx = wsx + P.gm.shift;
    if (hh >= P.gm.Hp) hh -= P.gm.Hp;
    if (wx >= P.gm.Wp) wx -= P.gm.Wp;
    tok = (hh < P.gm.H && wx < P.gm.W) ? (b * P.gm.H + hh) * P.gm.W + wx : -1;
    if (P.gm.shift > 0) {
      const int rh = hs < P.gm.Hp - 7 ? 0 : (hs < P.gm.Hp - P.gm.shift ? 1 : 2);
      const int rw = wsx < P.gm.Wp - 7 ? 0 : (wsx < P.gm.Wp - P.gm.shift ? 1 : 2);
      region = rh * 3 + rw;
    }
  }
  if (tok >= 0) {
    const __nv_bfloat16* qrow = P.qk + static_cast<int64_t>(tok) * 2 * C + h * HD;
    gather_row64(in_s, r, qrow);
    gather_row64(in_s + 8192, r, qrow + C);
    gather_row64(in_s + 16384, r, P.vb + static_cast<int64_t>(tok) * C + h * HD);
    if (BWD) gather_row64(in_s + 24576, r, P.dout + static_cast<int64_t>(tok) * C + h * HD);
  } else {
    zero_row64(in_g, r);
    zero_row64(in_g + 16384, r);
    if (BWD) zero_row64(in_g + 24576, r);
    if (tok == -1) bias_row64(in_g + 8192, r, P.qk_bias + C + h * HD);
    else zero_row64(in_g + 8192, r);
  }
  tok_out[r] = tok;
  wg_out[r] = wg;
  rid_out[r] = static_cast<uint8_t>(region);
}

// =====================================================================================================
// backward
// =====================================================================================================
__global__ void __launch_bounds__(kPipeThreads, 1)
attn_bwd_pipe_kernel(const AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* misc = gen + kBwdMisc;
  float* tbl = reinterpret_cast<float*>(misc + kMiscTbl);
  int* tok_s = reinterpret_cast<int*>(misc + kMiscTok);
  int* wg_s = reinterpret_cast<int*>(misc + kMiscWg);
  uint8_t* rid_s = misc + kMiscRid;
  const Bars bars{base + kBwdMisc + kMiscBar};
  const uint32_t tmem_ptr_addr = base + kBwdMisc + kMiscTmem;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(misc + kMiscTmem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int C = P.C;
  const int niter = (P.npairs - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 12) {
    if (lane == 0) init_bars(bars);
    __syncwarp();
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 176; i += kPipeThreads) tbl[i] = i < 169 ? __ldg(P.table + i * P.nH + h) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;

  if (warp >= 12) {
    // ================= MMA issuer =================
    reg_dealloc<56>();
    if (warp == 12 && lane == 0) {
      const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128, 128);  // S, dP : K-major x K-major
      const uint32_t idesc_t = make_idesc(1u, 1u, 1u, 128, HD);   // dV, dK: MN-major x MN-major
      const uint32_t idesc_q = make_idesc(1u, 0u, 1u, 128, HD);   // dQ    : K-major x MN-major
      auto second_stage = [&](int g2, int n2) {
        const uint32_t in = base + kBwdIn + (2 * g2) * 32768;
        const uint32_t Qs = in, Ks = in + 8192, Gs = in + 24576;
        const uint32_t Pb = base + bwd_p_tile(g2), Db = base + bwd_ds_tile(g2);
        const uint32_t t0 = tmem + g2 * 256;
        mbar_wait(bars.p_ready(g2), n2 & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // dV = Pbd^T dO : K = 128 query rows
          umma_bf16(t0, make_smem_desc(Pb + ks * 2048, 16384, 1024, kSwizzle128),
                    make_smem_desc(Gs + ks * 1024, 512, 512, kSwizzle64), idesc_t, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // dK = dSbd^T Q
          umma_bf16(t0 + 32, make_smem_desc(Db + ks * 2048, 16384, 1024, kSwizzle128),
                    make_smem_desc(Qs + ks * 1024, 512, 512, kSwizzle64), idesc_t, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // dQ = dSbd [K_A;K_B] : K = 128 stacked keys
          umma_bf16(t0 + 64, make_smem_desc(Db + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, kSwizzle128),
                    make_smem_desc(Ks + ks * 1024, 512, 512, kSwizzle64), idesc_q, ks > 0 ? 1u : 0u);
        umma_commit(bars.o_done(g2));        // results ready for the compute group
        umma_commit(bars.in_free(g2, 0));    // the lane's input tiles may be refilled
      };
      for (int i = 0; i < niter; ++i) {
        const int g = i & 1, n = i >> 1;
        const uint32_t in = base + kBwdIn + (2 * g) * 32768;
        const uint32_t Qs = in, Ks = in + 8192, Vs = in + 16384, Gs = in + 24576;
        const uint32_t t0 = tmem + g * 256;
        mbar_wait(bars.full(g, 0), n & 1);
        if (n > 0) mbar_wait(bars.t_free(g), (n - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)  // S = Q K^T -> cols [0,128)
          umma_bf16(t0, make_smem_desc(Qs + ks * 32, 16, 512, kSwizzle64), make_smem_desc(Ks + ks * 32, 16, 512, kSwizzle64),
                    idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)  // dP = dO V^T -> cols [128,256)
          umma_bf16(t0 + 128, make_smem_desc(Gs + ks * 32, 16, 512, kSwizzle64),
                    make_smem_desc(Vs + ks * 32, 16, 512, kSwizzle64), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bars.s_done(g));
        if (i >= 1) second_stage((i - 1) & 1, (i - 1) >> 1);
      }
      if (niter >= 1) second_stage((niter - 1) & 1, (niter - 1) >> 1);
    }
  } else if (warp >= 8) {
    // ================= loaders =================
    reg_dealloc<56>();
    const int r = threadIdx.x - 256;
    const int pi = (r & 63) / 7, pj = (r & 63) % 7;
    int pair = blockIdx.x;
    for (int i = 0; i < niter; ++i, pair += gridDim.x) {
      const int g = i & 1, n = i >> 1;
      const int b4 = 2 * g;  // backward: one input slot per lane (the other two slots hold lane 1's P / dS tiles)
      if (n >= 1) mbar_wait(bars.in_free(g, 0), (n - 1) & 1);  // the lane's previous pair is done with the tiles
      load_pair<true>(P, pair, r, pi, pj, h, base + kBwdIn + b4 * 32768, gen + kBwdIn + b4 * 32768, tok_s + b4 * 128,
                      wg_s + b4 * 128, rid_s + b4 * 128);
      cp_async_commit();
      if (i >= 1) {  // publish pair i-1 once its gathers have landed; pair i's stay in flight behind it
        cp_async_wait_group<1>();
        fence_proxy_async_smem();
        mbar_arrive(bars.full((i - 1) & 1, 0));
      }
    }
    if (niter >= 1) {
      cp_async_wait_group<0>();
      fence_proxy_async_smem();
      mbar_arrive(bars.full((niter - 1) & 1, 0));
    }
  } else {
    // ================= compute groups =================
    reg_alloc<200>();
    const int g = warp >> 2;
    const int r = threadIdx.x & 127;
    const int half = r >> 6, pos = r & 63;
    const uint32_t t0 = tmem + g * 256 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint8_t* Pb_g = gen + bwd_p_tile(g);
    uint8_t* Db_g = gen + bwd_ds_tile(g);
    const int bi = rpb_base(pos < kNTok ? pos : 0);
    const bool masked = P.gm.shift > 0;
    float dtab[kNTok];
#pragma unroll
    for (int j = 0; j < kNTok; ++j) dtab[j] = 0.f;

    for (int i = g, n = 0; i < niter; i += 2, ++n) {
      const int b4 = 2 * g;
      mbar_wait(bars.full(g, 0), n & 1);
      const int tok = tok_s[b4 * 128 + r];
      const int wg = wg_s[b4 * 128 + r];
      const uint8_t* rrow = rid_s + b4 * 128 + half * 64;
      const int my_region = rrow[pos];
      const float* xmask = (P.ext_mask != nullptr && tok != -2)
                               ? P.ext_mask + (static_cast<int64_t>(wg % P.ext_mask_nw) * kNTok + pos) * kNTok
                               : nullptr;
      mbar_wait(bars.s_done(g), n & 1);
      tc_fence_after();
      float p[64], ds[64];
      {
        uint32_t s0[32], s1[32];
        tmem_ld32(t0 + half * 64, s0);
        tmem_ld32(t0 + half * 64 + 32, s1);
        tmem_ld_wait();
        if (tok >= 0) {
          const float lse = __ldg(P.lse + (static_cast<int64_t>(wg) * P.nH + h) * 64 + pos);
#pragma unroll
          for (int j = 0; j < kNTok; ++j) {
            float s = __uint_as_float(j < 32 ? s0[j & 31] : s1[j & 31]) + tbl[bi - rpb_col(j)];
            if (masked && rrow[j] != my_region) s += -100.0f;
            if (xmask != nullptr) s += __ldg(xmask + j);
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
        tmem_ld32(t0 + 128 + half * 64, g0);
        tmem_ld32(t0 + 128 + half * 64 + 32, g1);
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
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t own = static_cast<uint32_t>(half) * 16384u + sw128_offset(r, c);
        const uint32_t oth = static_cast<uint32_t>(half ^ 1) * 16384u + sw128_offset(r, c);
        *reinterpret_cast<uint4*>(Pb_g + own) =
            make_uint4(pack_bf16(p[8 * c], p[8 * c + 1]), pack_bf16(p[8 * c + 2], p[8 * c + 3]),
                       pack_bf16(p[8 * c + 4], p[8 * c + 5]), pack_bf16(p[8 * c + 6], p[8 * c + 7]));
        *reinterpret_cast<uint4*>(Pb_g + oth) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Db_g + own) =
            make_uint4(pack_bf16(ds[8 * c], ds[8 * c + 1]), pack_bf16(ds[8 * c + 2], ds[8 * c + 3]),
                       pack_bf16(ds[8 * c + 4], ds[8 * c + 5]), pack_bf16(ds[8 * c + 6], ds[8 * c + 7]));
        *reinterpret_cast<uint4*>(Db_g + oth) = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bars.p_ready(g));

      mbar_wait(bars.o_done(g), n & 1);
      tc_fence_after();
      uint32_t a[32], b[32], c2[32];
      tmem_ld32(t0, a);
      tmem_ld32(t0 + 32, b);
      tmem_ld32(t0 + 64, c2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bars.t_free(g));  // TMEM columns of this lane may be overwritten by the next S/dP
      if (tok >= 0) {
        float4* dvp = reinterpret_cast<float4*>(P.dv + static_cast<int64_t>(tok) * C + h * HD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v = make_float4(__uint_as_float(a[4 * c]), __uint_as_float(a[4 * c + 1]), __uint_as_float(a[4 * c + 2]),
                                 __uint_as_float(a[4 * c + 3]));
          if (P.dv_acc) {
            const float4 o = dvp[c];
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          dvp[c] = v;
        }
        uint4* dk = reinterpret_cast<uint4*>(P.dqk + static_cast<int64_t>(tok) * 2 * C + C + h * HD);
        uint4* dq = reinterpret_cast<uint4*>(P.dqk + static_cast<int64_t>(tok) * 2 * C + h * HD);
        const float sc = P.scale;  // d(xW+b) = dq * scale because q was stored pre-scaled
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          dk[c] = make_uint4(pack_bf16(__uint_as_float(b[8 * c]), __uint_as_float(b[8 * c + 1])),
                             pack_bf16(__uint_as_float(b[8 * c + 2]), __uint_as_float(b[8 * c + 3])),
                             pack_bf16(__uint_as_float(b[8 * c + 4]), __uint_as_float(b[8 * c + 5])),
                             pack_bf16(__uint_as_float(b[8 * c + 6]), __uint_as_float(b[8 * c + 7])));
          dq[c] = make_uint4(pack_bf16(sc * __uint_as_float(c2[8 * c]), sc * __uint_as_float(c2[8 * c + 1])),
                             pack_bf16(sc * __uint_as_float(c2[8 * c + 2]), sc * __uint_as_float(c2[8 * c + 3])),
                             pack_bf16(sc * __uint_as_float(c2[8 * c + 4]), sc * __uint_as_float(c2[8 * c + 5])),
                             pack_bf16(sc * __uint_as_float(c2[8 * c + 6]), sc * __uint_as_float(c2[8 * c + 7])));
        }
      }
      // zero-padded keys: k == bias, so their gradient goes to the k half of qk.bias (warp-reduced, one atomic/lane)
      if (__any_sync(0xffffffffu, tok == -1)) {
        float mine = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = warp_sum(tok == -1 ? __uint_as_float(b[j]) : 0.f);
          if ((threadIdx.x & 31) == j) mine = s;
        }
        atomicAdd(P.d_qk_bias + C + h * HD + (threadIdx.x & 31), mine);
      }
    }

    // relative-position-bias gradient: dTable[idx(i,j), h] += sum over my pairs of dS[i][j]
    named_bar_sync(1, 256);  // both compute groups are past their last read of tbl
    if (threadIdx.x < 176) tbl[threadIdx.x] = 0.f;
    named_bar_sync(1, 256);
    if (pos < kNTok) {
#pragma unroll
      for (int j = 0; j < kNTok; ++j) atomicAdd(tbl + bi - rpb_col(j), dtab[j]);
    }
    named_bar_sync(1, 256);
    if (threadIdx.x < 169) atomicAdd(P.d_table + threadIdx.x * P.nH + h, tbl[threadIdx.x]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
__global__ void __launch_bounds__(kPipeThreads, 1)
attn_fwd_pipe_kernel(const AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* misc = gen + kFwdMisc;
  float* tbl = reinterpret_cast<float*>(misc + kMiscTbl);
  int* tok_s = reinterpret_cast<int*>(misc + kMiscTok);
  int* wg_s = reinterpret_cast<int*>(misc + kMiscWg);
  uint8_t* rid_s = misc + kMiscRid;
  const Bars bars{base + kFwdMisc + kMiscBar};
  const uint32_t tmem_ptr_addr = base + kFwdMisc + kMiscTmem;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(misc + kMiscTmem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int C = P.C;
  const int niter = (P.npairs - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 12) {
    if (lane == 0) init_bars(bars);
    __syncwarp();
    tmem_alloc(tmem_ptr_addr, 256);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 176; i += kPipeThreads) tbl[i] = i < 169 ? __ldg(P.table + i * P.nH + h) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;

  if (warp >= 12) {
    reg_dealloc<56>();
    if (warp == 12 && lane == 0) {
      const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128, 128);
      const uint32_t idesc_o = make_idesc(1u, 0u, 1u, 128, HD);
      auto second_stage = [&](int g2, int n2) {  // O = P V: window A -> cols [0,hd), window B -> [hd,2hd)
        const uint32_t Vs = base + kFwdIn + (2 * g2 + (n2 & 1)) * 32768 + 16384;
        const uint32_t Ps = base + kFwdTiles + g2 * 16384;
        const uint32_t t0 = tmem + g2 * 128;
        mbar_wait(bars.p_ready(g2), n2 & 1);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(t0 + half * HD, make_smem_desc(Ps + ks * 32, 16, 1024, kSwizzle128),
                      make_smem_desc(Vs + half * 4096 + ks * 1024, 512, 512, kSwizzle64), idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(bars.o_done(g2));
        umma_commit(bars.in_free(g2, n2 & 1));
      };
      for (int i = 0; i < niter; ++i) {
        const int g = i & 1, n = i >> 1;
        const uint32_t Qs = base + kFwdIn + (2 * g + (n & 1)) * 32768, Ks = Qs + 8192;
        mbar_wait(bars.full(g, n & 1), (n >> 1) & 1);
        if (n > 0) mbar_wait(bars.t_free(g), (n - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tmem + g * 128, make_smem_desc(Qs + ks * 32, 16, 512, kSwizzle64),
                    make_smem_desc(Ks + ks * 32, 16, 512, kSwizzle64), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bars.s_done(g));
        if (i >= 1) second_stage((i - 1) & 1, (i - 1) >> 1);
      }
      if (niter >= 1) second_stage((niter - 1) & 1, (niter - 1) >> 1);
    }
  } else if (warp >= 8) {
    reg_dealloc<56>();
    const int r = threadIdx.x - 256;
    const int pi = (r & 63) / 7, pj = (r & 63) % 7;
    int pair = blockIdx.x;
    for (int i = 0; i < niter; ++i, pair += gridDim.x) {
      const int g = i & 1, n = i >> 1;
      const int slot = n & 1, b4 = 2 * g + slot;
      if (n >= 2) mbar_wait(bars.in_free(g, slot), ((n >> 1) - 1) & 1);
      load_pair<false>(P, pair, r, pi, pj, h, base + kFwdIn + b4 * 32768, gen + kFwdIn + b4 * 32768, tok_s + b4 * 128,
                       wg_s + b4 * 128, rid_s + b4 * 128);
      cp_async_commit();
      if (i >= 2) {  // three pairs of gathers in flight: publish pair i-2 once it has landed
        cp_async_wait_group<2>();
        fence_proxy_async_smem();
        mbar_arrive(bars.full(i & 1, ((i - 2) >> 1) & 1));
      }
    }
    for (int j = niter >= 2 ? niter - 2 : 0; j < niter; ++j) {  // drain
      if (j == niter - 1) cp_async_wait_group<0>(); else cp_async_wait_group<1>();
      fence_proxy_async_smem();
      mbar_arrive(bars.full(j & 1, (j >> 1) & 1));
    }
  } else {
    reg_alloc<200>();
    const int g = warp >> 2;
    const int r = threadIdx.x & 127;
    const int half = r >> 6, pos = r & 63;
    const uint32_t t0 = tmem + g * 128 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint8_t* Ps_g = gen + kFwdTiles + g * 16384;
    const int bi = rpb_base(pos < kNTok ? pos : 0);
    const bool masked = P.gm.shift > 0;

    for (int i = g, n = 0; i < niter; i += 2, ++n) {
      const int b4 = 2 * g + (n & 1);
      mbar_wait(bars.full(g, n & 1), (n >> 1) & 1);
      const int tok = tok_s[b4 * 128 + r];
      const int wg = wg_s[b4 * 128 + r];
      const uint8_t* rrow = rid_s + b4 * 128 + half * 64;
      const int my_region = rrow[pos];
      const float* xmask = (P.ext_mask != nullptr && tok != -2)
                               ? P.ext_mask + (static_cast<int64_t>(wg % P.ext_mask_nw) * kNTok + pos) * kNTok
                               : nullptr;
      mbar_wait(bars.s_done(g), n & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(t0 + half * 64, s0);
      tmem_ld32(t0 + half * 64 + 32, s1);
      tmem_ld_wait();
      float p[64];
      if (tok != -2) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kNTok; ++j) {
          float s = __uint_as_float(j < 32 ? s0[j & 31] : s1[j & 31]) + tbl[bi - rpb_col(j)];
          if (masked && rrow[j] != my_region) s += -100.0f;
          if (xmask != nullptr) s += __ldg(xmask + j);
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
        *reinterpret_cast<uint4*>(Ps_g + sw128_offset(r, c)) =
            make_uint4(pack_bf16(p[8 * c], p[8 * c + 1]), pack_bf16(p[8 * c + 2], p[8 * c + 3]),
                       pack_bf16(p[8 * c + 4], p[8 * c + 5]), pack_bf16(p[8 * c + 6], p[8 * c + 7]));
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bars.p_ready(g));

      mbar_wait(bars.o_done(g), n & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(t0 + half * HD, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bars.t_free(g));
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace

int launch_attn_fwd_pipe(const AttnParams& P, const crf_block_desc& d, cudaStream_t st) {
  const size_t smem = kFwdMisc + kMiscBytes + 1024;
  CRF_CUDA(cudaFuncSetAttribute(attn_fwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int gx = num_sms(d.device) / P.nH;   // one 512-thread CTA per SM, never a partial second wave
  if (gx > (P.npairs + 1) / 2) gx = (P.npairs + 1) / 2;  // at least two pairs per CTA so both lanes work
  if (gx < 1) gx = 1;
  const double TC = static_cast<double>(d.B) * d.H * d.W * d.C;
  KernelTimer tm(st, 4.0 * 49 * 49 * d.C * P.total_windows, 8.0 * TC, "attn_fwd_B%d_%dx%d_C%d_s%d", d.B, d.H, d.W, d.C,
                 d.shift);
  attn_fwd_pipe_kernel<<<dim3(gx, P.nH), kPipeThreads, smem, st>>>(P);
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_attn_bwd_pipe(const AttnParams& P, const crf_block_desc& d, cudaStream_t st) {
  const size_t smem = kBwdMisc + kMiscBytes + 1024;
  CRF_CUDA(cudaFuncSetAttribute(attn_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int gx = num_sms(d.device) / P.nH;
  if (gx > (P.npairs + 1) / 2) gx = (P.npairs + 1) / 2;
  if (gx < 1) gx = 1;
  const double TC = static_cast<double>(d.B) * d.H * d.W * d.C;
  KernelTimer tm(st, 10.0 * 49 * 49 * d.C * P.total_windows, 16.0 * TC, "attn_bwd_B%d_%dx%d_C%d_s%d", d.B, d.H, d.W,
                 d.C, d.shift);
  attn_bwd_pipe_kernel<<<dim3(gx, P.nH), kPipeThreads, smem, st>>>(P);
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace crf
