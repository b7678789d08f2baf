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

// t = number of the step being taken (1 for the first).  The hyper-parameters are Python doubles in torch.optim.Adam and
// torch derives 1 - beta, the bias corrections and lr / bc1 in double before rounding once to the parameter's precision
// (with fp32 arithmetic 1 - 0.999f is 0.0009999871, a 1.3e-5 relative bias on every exp_avg_sq): same here.
CRF_HD AdamCoef adam_coef(double lr, double b1, double b2, double eps, double wd, double t) {
  AdamCoef c;
  c.b1 = static_cast<float>(b1);
  c.b2 = static_cast<float>(b2);
  c.one_minus_b1 = static_cast<float>(1.0 - b1);
  c.one_minus_b2 = static_cast<float>(1.0 - b2);
  const double bc1 = 1.0 - pow(b1, t);
  const double bc2 = 1.0 - pow(b2, t);
  c.step_size = static_cast<float>(lr / bc1);
  c.sqrt_bc2 = static_cast<float>(sqrt(bc2));
  c.eps = static_cast<float>(eps);
  c.wd = static_cast<float>(wd);
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
