// The training loop's loss in two kernels (/root/reference/src/train.py:94-100):
//   loss = 1.0 * SSIM(pred, target) + 0.1 * L1(pred, target)
// with the monodepth2-style SSIM of /root/reference/src/loss.py:57-88: reflection pad 1, 3x3 average pools of
// x, y, x^2, y^2, xy, s = clamp((1 - n / d) / 2, 0, 1), mean over all pixels.
//
// Stock PyTorch evaluates this as ~40 small kernels per step (pads, five pools, a dozen elementwise passes and their
// backwards).  Here the forward kernel computes, per pixel p, the five window moments from nine reflected taps, the
// SSIM value and the three partial derivatives of the loss with respect to the window moments that depend on pred
//   G1 = dL/d mean(x),  G2 = dL/d mean(x^2),  G3 = dL/d mean(xy)       (0 where the clamp is active)
// and block-reduces the two sums.  The backward kernel applies the adjoint of (reflection pad o 3x3 box filter):
//   dL/dx(q) = 1/9 * sum over padded positions u that read q [ T1(u) + 2 x(q) T2(u) + y(q) T3(u) ] + 0.1 sign(x-y)/N,
//   Tk(u) = sum of Gk over the (valid) window centres p within distance 1 of u.
#include "crf_host.h"
#include "crf_ptx.cuh"

namespace crf {

namespace {

constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

__device__ __forceinline__ float ldv(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldv(const __nv_bfloat16* p) {
  return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
__device__ __forceinline__ void stv(float* p, float v) { *p = v; }
__device__ __forceinline__ void stv(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

template <typename TP>
__global__ void __launch_bounds__(256)
loss_fwd_kernel(const TP* __restrict__ pred, const float* __restrict__ tgt, int H, int W, float* __restrict__ sums,
                float* __restrict__ G, int64_t plane_stride_g) {
  pdl_prologue();
  const int w = blockIdx.x * 32 + threadIdx.x;
  const int64_t img = static_cast<int64_t>(blockIdx.z) * H * W;
  float v_ssim = 0.f, v_l1 = 0.f;
  // a block walks down the image in 8-row tiles: a few hundred blocks (not ~10 k) end up adding to the two sums
  for (int h = blockIdx.y * 8 + threadIdx.y; h < H; h += gridDim.y * 8) {
    if (w >= W) break;
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f, x0 = 0.f, y0 = 0.f;
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      const int hh = reflect1(h + dh, H);
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        const int ww = reflect1(w + dw, W);
        const float x = ldv(pred + img + static_cast<int64_t>(hh) * W + ww);
        const float y = __ldg(tgt + img + static_cast<int64_t>(hh) * W + ww);
        sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
        if (dh == 0 && dw == 0) { x0 = x; y0 = y; }
      }
    }
    const float k9 = 1.0f / 9.0f;
    const float mx = sx * k9, my = sy * k9;
    const float vx = sxx * k9 - mx * mx, vy = syy * k9 - my * my, cxy = sxy * k9 - mx * my;
    const float A = 2.f * mx * my + kC1, Bq = 2.f * cxy + kC2;
    const float Cq = mx * mx + my * my + kC1, Dq = vx + vy + kC2;
    const float inv_cd = 1.0f / (Cq * Dq);
    const float r = A * Bq * inv_cd;
    const float s = 0.5f * (1.0f - r);
    v_ssim += fminf(fmaxf(s, 0.f), 1.f);
    v_l1 += fabsf(x0 - y0);
    if (G != nullptr) {
      // dr/d(mean x), dr/d(mean x^2), dr/d(mean xy); ds = -dr / 2; the clamp passes gradients on [0, 1] only
      const float pass = (s >= 0.f && s <= 1.f) ? -0.5f : 0.f;
      const float dr_dmx = 2.f * my * (Bq - A) * inv_cd - r * (2.f * mx / Cq - 2.f * mx / Dq);
      const float dr_dm2 = -r / Dq;
      const float dr_dm3 = 2.f * A * inv_cd;
      const int64_t o = img + static_cast<int64_t>(h) * W + w;
      G[o] = pass * dr_dmx;
      G[o + plane_stride_g] = pass * dr_dm2;
      G[o + 2 * plane_stride_g] = pass * dr_dm3;
    }
  }
  __shared__ float red[2][8];
  v_ssim = warp_sum(v_ssim);
  v_l1 = warp_sum(v_l1);
  if (threadIdx.x == 0) { red[0][threadIdx.y] = v_ssim; red[1][threadIdx.y] = v_l1; }
  __syncthreads();
  if (threadIdx.y == 0 && threadIdx.x < 2) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
    atomicAdd(sums + threadIdx.x, t);
  }
}

template <typename TP>
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const TP* __restrict__ pred, const float* __restrict__ tgt, const float* __restrict__ G,
                int64_t plane_stride_g, const float* __restrict__ gout, int H, int W, float inv_n,
                TP* __restrict__ dpred) {
  pdl_prologue();
  const int w = blockIdx.x * 32 + threadIdx.x, h = blockIdx.y * 8 + threadIdx.y;
  if (h >= H || w >= W) return;
  const int64_t img = static_cast<int64_t>(blockIdx.z) * H * W;
  const int64_t o = img + static_cast<int64_t>(h) * W + w;
  const float x = ldv(pred + o), y = __ldg(tgt + o);
  // padded positions that read q: q itself and its mirror images across the border
  int uh[3], uw[3], nh = 1, nw = 1;
  uh[0] = h; uw[0] = w;
  if (h == 1) uh[nh++] = -1;
  if (h == H - 2) uh[nh++] = H;
  if (w == 1) uw[nw++] = -1;
  if (w == W - 2) uw[nw++] = W;
  float t1 = 0.f, t2 = 0.f, t3 = 0.f;
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b)
#pragma unroll
      for (int dh = -1; dh <= 1; ++dh) {
        const int ph = uh[a] + dh;
        if (ph < 0 || ph >= H) continue;
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
          const int pw = uw[b] + dw;
          if (pw < 0 || pw >= W) continue;
          const int64_t po = img + static_cast<int64_t>(ph) * W + pw;
          t1 += __ldg(G + po);
          t2 += __ldg(G + po + plane_stride_g);
          t3 += __ldg(G + po + 2 * plane_stride_g);
        }
      }
  const float d = x - y;
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  const float g = __ldg(gout) * inv_n * ((t1 + 2.f * x * t2 + y * t3) * (1.0f / 9.0f) + 0.1f * sgn);
  stv(dpred + o, g);
}

}  // namespace

int launch_depth_loss_fwd(const void* pred, int pred_dtype, const float* tgt, int n_img, int H, int W, float* sums,
                          float* G, cudaStream_t st) {
  CRF_CHECK(H >= 2 && W >= 2 && n_img > 0, "depth_loss: needs H, W >= 2 (reflection pad 1), got %dx%d", H, W);
  CRF_CHECK(pred_dtype == CRF_DT_F32 || pred_dtype == CRF_DT_BF16, "depth_loss: unsupported dtype %d", pred_dtype);
  const int row_tiles = (H + 7) / 8;
  const dim3 grid((W + 31) / 32, row_tiles < 4 ? row_tiles : 4, n_img), block(32, 8);
  const int64_t plane = static_cast<int64_t>(n_img) * H * W;
  KernelTimer tm(st, 0.0, static_cast<double>(plane) * (8 + (G != nullptr ? 12 : 0)), "depth_loss_fwd_%dx%dx%d", n_img, H, W);
  if (pred_dtype == CRF_DT_F32)
    launch_pdl((loss_fwd_kernel<float>), grid, block, 0, st, reinterpret_cast<const float*>(pred), tgt, H, W, sums, G, plane);
  else
    launch_pdl((loss_fwd_kernel<__nv_bfloat16>), grid, block, 0, st, reinterpret_cast<const __nv_bfloat16*>(pred), tgt, H, W, sums,
                                                           G, plane);
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_depth_loss_bwd(const void* pred, int pred_dtype, const float* tgt, const float* G, const float* gout,
                          int n_img, int H, int W, void* dpred, cudaStream_t st) {
  CRF_CHECK(H >= 2 && W >= 2 && n_img > 0, "depth_loss: needs H, W >= 2 (reflection pad 1), got %dx%d", H, W);
  CRF_CHECK(pred_dtype == CRF_DT_F32 || pred_dtype == CRF_DT_BF16, "depth_loss: unsupported dtype %d", pred_dtype);
  const dim3 grid((W + 31) / 32, (H + 7) / 8, n_img), block(32, 8);
  const int64_t plane = static_cast<int64_t>(n_img) * H * W;
  const float inv_n = 1.0f / static_cast<float>(plane);
  KernelTimer tm(st, 0.0, static_cast<double>(plane) * 24, "depth_loss_bwd_%dx%dx%d", n_img, H, W);
  if (pred_dtype == CRF_DT_F32)
    launch_pdl((loss_bwd_kernel<float>), grid, block, 0, st, reinterpret_cast<const float*>(pred), tgt, G, plane, gout, H, W, inv_n,
                                                   reinterpret_cast<float*>(dpred));
  else
    launch_pdl((loss_bwd_kernel<__nv_bfloat16>), grid, block, 0, st, reinterpret_cast<const __nv_bfloat16*>(pred), tgt, G, plane,
                                                           gout, H, W, inv_n, reinterpret_cast<__nv_bfloat16*>(dpred));
  CRF_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace crf
