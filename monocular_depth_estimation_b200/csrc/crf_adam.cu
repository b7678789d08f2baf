// Multi-tensor Adam step for the reference loop's optimizer (src/train.py:41, :108: torch.optim.Adam over all model
// parameters).  HBM-bound: 16 B read + 12 B written per parameter.  The caller hands over a HOST
// array of (p, g, m, v, n) records; they travel to the device as kernel arguments, 80 tensors per launch (no device
// table, nothing to keep alive or in sync, trivially capturable in a CUDA graph: ~300 parameter tensors = 4 launches).
// A CTA finds its (tensor, chunk) from the per-batch prefix of chunk counts and streams its chunk with 16-byte vector
// accesses (scalar when a view is not 16-byte aligned).  The step counter lives in device memory (one float,
// incremented by a last tiny launch), so replays of a captured step keep counting.
//
// STATUS: the element update (crf_adam_math.h) is verified on the host against torch.optim.Adam; the kernel on a B200
// against the same update through tools/hwcheck (profiles/r01_hwcheck.txt: bit-identical).  Speed and graph capture are
// unmeasured: opt-in (training.LibAdam / bench.py --lib-adam).
#include "crf_adam_math.h"
#include "crf_host.h"

namespace crf {

namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamBatch = 80;   // tensors per launch: 80 x 44 B of kernel arguments (limit 4 KB)

struct AdamBatch {
  float* p[kAdamBatch];
  const float* g[kAdamBatch];
  float* m[kAdamBatch];
  float* v[kAdamBatch];
  long long n[kAdamBatch];
  int chunk_end[kAdamBatch];  // running total of chunks up to and including tensor i
  int count;
};
static_assert(sizeof(AdamBatch) <= 4000, "kernel-argument budget");

__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(const __grid_constant__ AdamBatch B, int chunk_elems, double lr, double b1, double b2, double eps,
                  double wd, const float* __restrict__ step) {
  int ti = 0;
  while (ti + 1 < B.count && static_cast<int>(blockIdx.x) >= B.chunk_end[ti]) ++ti;   // uniform per CTA, <= 80 steps
  const int chunk = static_cast<int>(blockIdx.x) - (ti > 0 ? B.chunk_end[ti - 1] : 0);
  __shared__ AdamCoef c_sh;   // double-precision pow / sqrt once per CTA, not once per thread
  if (threadIdx.x == 0) c_sh = adam_coef(lr, b1, b2, eps, wd, static_cast<double>(*step) + 1.0);
  __syncthreads();
  const AdamCoef c = c_sh;
  const long long lo = static_cast<long long>(chunk) * chunk_elems;
  long long hi = lo + chunk_elems;
  if (hi > B.n[ti]) hi = B.n[ti];
  float* p = B.p[ti] + lo;
  const float* g = B.g[ti] + lo;
  float* m = B.m[ti] + lo;
  float* v = B.v[ti] + lo;
  const int n = static_cast<int>(hi - lo);
  // p, m, v vectorised when 16-byte aligned (separately allocated parameters and the optimizer's padded state views
  // always are); the gradient may be an arbitrarily aligned view of a DDP bucket: four scalar loads then
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const bool gvec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  int done = 0;
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += kAdamThreads) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      float4 gg;
      if (gvec) gg = __ldg(reinterpret_cast<const float4*>(g) + i);
      else gg = make_float4(__ldg(g + 4 * i), __ldg(g + 4 * i + 1), __ldg(g + 4 * i + 2), __ldg(g + 4 * i + 3));
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_update(c, pp.x, gg.x, mm.x, vv.x);
      adam_update(c, pp.y, gg.y, mm.y, vv.y);
      adam_update(c, pp.z, gg.z, mm.z, vv.z);
      adam_update(c, pp.w, gg.w, mm.w, vv.w);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < n; i += kAdamThreads) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_update(c, pp, g[i], mm, vv);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

__global__ void adam_count_kernel(float* step) { *step += 1.0f; }

}  // namespace

int launch_adam_step(const crf_adam_tensor* tensors, int n_tensors, int chunk_elems, double lr, double b1, double b2,
                     double eps, double wd, float* step, cudaStream_t st) {
  CRF_CHECK(tensors != nullptr && step != nullptr, "crf_adam_step: null pointer");
  CRF_CHECK(n_tensors > 0, "crf_adam_step: no tensors");
  CRF_CHECK(chunk_elems >= 1024 && chunk_elems % 4 == 0, "crf_adam_step: chunk_elems=%d must be a multiple of 4, >= 1024",
            chunk_elems);
  CRF_CHECK(b1 >= 0. && b1 < 1. && b2 >= 0. && b2 < 1. && eps >= 0. && lr >= 0. && wd >= 0.,
            "crf_adam_step: bad hyper-parameters");
  double total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    CRF_CHECK(tensors[i].p && tensors[i].g && tensors[i].m && tensors[i].v && tensors[i].n > 0,
              "crf_adam_step: tensor %d has a null pointer or no elements", i);
    CRF_CHECK(((reinterpret_cast<uintptr_t>(tensors[i].p) | reinterpret_cast<uintptr_t>(tensors[i].g) |
                reinterpret_cast<uintptr_t>(tensors[i].m) | reinterpret_cast<uintptr_t>(tensors[i].v)) & 3) == 0,
              "crf_adam_step: tensor %d is not 4-byte aligned", i);
    total += static_cast<double>(tensors[i].n);
  }
  KernelTimer tm(st, 0.0, 28.0 * total, "adam_step_%dtensors", n_tensors);
  int launches = 0;
  for (int i0 = 0; i0 < n_tensors; i0 += kAdamBatch) {
    AdamBatch B;
    B.count = n_tensors - i0 < kAdamBatch ? n_tensors - i0 : kAdamBatch;
    long long chunks = 0;
    for (int k = 0; k < B.count; ++k) {
      const crf_adam_tensor& t = tensors[i0 + k];
      B.p[k] = t.p;
      B.g[k] = t.g;
      B.m[k] = t.m;
      B.v[k] = t.v;
      B.n[k] = t.n;
      chunks += (t.n + chunk_elems - 1) / chunk_elems;
      CRF_CHECK(chunks < (1LL << 30), "crf_adam_step: too many chunks");
      B.chunk_end[k] = static_cast<int>(chunks);
    }
    for (int k = B.count; k < kAdamBatch; ++k) {
      B.p[k] = nullptr; B.g[k] = nullptr; B.m[k] = nullptr; B.v[k] = nullptr; B.n[k] = 0;
      B.chunk_end[k] = static_cast<int>(chunks);
    }
    adam_multi_kernel<<<static_cast<unsigned>(chunks), kAdamThreads, 0, st>>>(B, chunk_elems, lr, b1, b2, eps, wd, step);
    CRF_CUDA(cudaGetLastError());
    ++launches;
  }
  adam_count_kernel<<<1, 1, 0, st>>>(step);
  CRF_CUDA(cudaGetLastError());
  note_launch(launches + 1);
  return 0;
}

}  // namespace crf
