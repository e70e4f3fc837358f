// optim.cu -- fused RAdam update over a flat fp32 span ("next" row 1).
//
// Replaces the per-tensor chain of ~10 ATen ops in RAdam.step (reference radam.py:34-92): second/first
// moment update (:58-59), optional decoupled weight decay (:82-83), adaptive step (:84-85) or the
// degenerated-SGD step (:88-90).  The rectification term N_sma and step_size (:62-78) depend only on the
// step count and are computed on the host exactly as the reference does; `mode` tells the kernel which
// branch applies.  One pass reads p, g, m, v and writes p, m, v (28 bytes per parameter).
#include "common.cuh"

namespace hn {

// mode: 0 = moments only (N_sma < 5, not degenerated: parameters untouched), 1 = adaptive, 2 = SGD-like
template <int MODE>
__global__ void __launch_bounds__(256)
radam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, float beta1, float beta2, float eps, float wd_lr, float step_lr, float grad_scale) {
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = __ldg(g + i) * grad_scale;
    // exp_avg_sq.mul_(beta2).addcmul_(1 - beta2, grad, grad)   (:58)
    const float vi = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(__fmul_rn(omb2, gi), gi));
    // exp_avg.mul_(beta1).add_(1 - beta1, grad)                (:59)
    const float mi = __fadd_rn(__fmul_rn(m[i], beta1), __fmul_rn(omb1, gi));
    v[i] = vi;
    m[i] = mi;
    if (MODE != 0) {
      float pi = p[i];
      if (wd_lr != 0.f) pi = __fadd_rn(pi, __fmul_rn(-wd_lr, pi));              // :83 / :89
      if (MODE == 1) {
        const float denom = __fadd_rn(__fsqrt_rn(vi), eps);                       // :84
        pi = __fadd_rn(pi, __fmul_rn(-step_lr, __fdiv_rn(mi, denom)));            // :85 addcdiv
      } else {
        pi = __fadd_rn(pi, __fmul_rn(-step_lr, mi));                              // :90
      }
      p[i] = pi;
    }
  }
}

// Same update with the step-dependent scalars read from device memory, so that the launch can live inside
// a CUDA graph: hp = {beta1, beta2, eps, wd*lr, step_size*lr, grad_scale, mode, -}.
__global__ void __launch_bounds__(256)
radam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, const float* __restrict__ hp) {
  const float beta1 = __ldg(hp), beta2 = __ldg(hp + 1), eps = __ldg(hp + 2), wd_lr = __ldg(hp + 3),
              step_lr = __ldg(hp + 4), grad_scale = __ldg(hp + 5);
  const int mode = (int)__ldg(hp + 6);
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = __ldg(g + i) * grad_scale;
    const float vi = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(__fmul_rn(omb2, gi), gi));
    const float mi = __fadd_rn(__fmul_rn(m[i], beta1), __fmul_rn(omb1, gi));
    v[i] = vi;
    m[i] = mi;
    if (mode != 0) {
      float pi = p[i];
      if (wd_lr != 0.f) pi = __fadd_rn(pi, __fmul_rn(-wd_lr, pi));
      if (mode == 1) {
        const float denom = __fadd_rn(__fsqrt_rn(vi), eps);
        pi = __fadd_rn(pi, __fmul_rn(-step_lr, __fdiv_rn(mi, denom)));
      } else {
        pi = __fadd_rn(pi, __fmul_rn(-step_lr, mi));
      }
      p[i] = pi;
    }
  }
}

}  // namespace hn

extern "C" int hn_radam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hp,
                                 void* stream) {
  HN_REQUIRE(n >= 0, "hn_radam_step_dev: negative n");
  if (n == 0) return 0;
  HN_REQUIRE(p && g && m && v && hp, "hn_radam_step_dev: null pointer");
  const int64_t want = (n + 255) / 256;
  const int64_t cap = (int64_t)hn::sm_count() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  hn::radam_dev_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hp);
  return hn::check_launch("radam_dev_kernel");
}

extern "C" int hn_radam_step(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2,
                             float eps, float lr, float weight_decay, float step_size, int mode, float grad_scale,
                             void* stream) {
  HN_REQUIRE(n >= 0, "hn_radam_step: negative n");
  HN_REQUIRE(mode >= 0 && mode <= 2, "hn_radam_step: mode must be 0, 1 or 2");
  if (n == 0) return 0;
  HN_REQUIRE(p && g && m && v, "hn_radam_step: null pointer");
  const int64_t want = (n + 255) / 256;
  const int64_t cap = (int64_t)hn::sm_count() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaStream_t s = (cudaStream_t)stream;
  // the reference forms the scalars in double and hands them to ATen as python floats (:83, :85)
  const float wd_lr = (float)((double)weight_decay * (double)lr);
  const float step_lr = (float)((double)step_size * (double)lr);
  switch (mode) {
    case 0: hn::radam_kernel<0><<<grid, 256, 0, s>>>(p, g, m, v, n, beta1, beta2, eps, wd_lr, step_lr, grad_scale); break;
    case 1: hn::radam_kernel<1><<<grid, 256, 0, s>>>(p, g, m, v, n, beta1, beta2, eps, wd_lr, step_lr, grad_scale); break;
    default: hn::radam_kernel<2><<<grid, 256, 0, s>>>(p, g, m, v, n, beta1, beta2, eps, wd_lr, step_lr, grad_scale); break;
  }
  return hn::check_launch("radam_kernel");
}
