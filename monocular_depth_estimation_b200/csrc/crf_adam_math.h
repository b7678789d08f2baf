// The Adam update of one element, shared by the device kernel (crf_adam.cu) and the host-side unit test
// (tests/test_adam_host.py compiles this header with g++): the update of torch.optim.Adam as the reference loop uses it
// (src/train.py:41: Adam(model.parameters(), lr), no amsgrad; L2 weight decay as torch applies it: g += wd * p).
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CRF_HD __host__ __device__ __forceinline__
#else
#define CRF_HD inline
#endif

namespace crf {

struct AdamCoef {
  float b1, b2, one_minus_b1, one_minus_b2, step_size, sqrt_bc2, eps, wd;
};

// t = number of the step being taken (1 for the first).
CRF_HD AdamCoef adam_coef(float lr, float b1, float b2, float eps, float wd, float t) {
  AdamCoef c;
  c.b1 = b1;
  c.b2 = b2;
  c.one_minus_b1 = 1.0f - b1;
  c.one_minus_b2 = 1.0f - b2;
  const float bc1 = 1.0f - powf(b1, t);
  const float bc2 = 1.0f - powf(b2, t);
  c.step_size = lr / bc1;
  c.sqrt_bc2 = sqrtf(bc2);
  c.eps = eps;
  c.wd = wd;
  return c;
}

CRF_HD void adam_update(const AdamCoef& c, float& p, float g, float& m, float& v) {
  if (c.wd != 0.0f) g = fmaf(c.wd, p, g);
  m = fmaf(c.one_minus_b1, g - m, m);          // lerp(m, g, 1 - b1), as torch's fused kernel does
  v = fmaf(c.b2, v, c.one_minus_b2 * g * g);
  const float denom = sqrtf(v) / c.sqrt_bc2 + c.eps;   // same operation order as torch's fused Adam kernel
  p -= c.step_size * m / denom;
}

}  // namespace crf
