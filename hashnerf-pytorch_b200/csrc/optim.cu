// optim.cu -- fused RAdam update over a flat fp32 span ("next" row 1).
//
// Replaces the per-tensor chain of ~10 ATen ops in RAdam.step (reference radam.py:34-92): second/first
// moment update (:58-59), optional decoupled weight decay (:82-83), adaptive step (:84-85) or the
// degenerated-SGD step (:88-90).  The rectification term N_sma and step_size (:62-78) depend only on the
// step count and are computed on the host exactly as the reference does; `mode` tells the kernel which
// branch applies.  One pass reads p, g, m, v and writes p, m, v (28 bytes per parameter) and, on request,
// clears g on the way (run_nerf.py:612 optimizer.zero_grad() folded in: +4 bytes instead of a separate fill).
#include "common.cuh"

namespace hn {

struct RadamScalars {
  float beta1, beta2, eps, wd_lr, step_lr, grad_scale;
  int mode;       // 0 = moments only (N_sma < 5, not degenerated: parameters untouched), 1 = adaptive, 2 = SGD-like
  int zero_grad;  // clear the gradient in the same pass (optimizer.zero_grad() folded in, run_nerf.py:612)
};

// one element, every operation an individually rounded fp32 op in the reference's order
__device__ __forceinline__ void radam_one(float& p, float g, float& m, float& v, const RadamScalars& s, float omb1,
                                          float omb2) {
  const float gi = g * s.grad_scale;
  // exp_avg_sq.mul_(beta2).addcmul_(1 - beta2, grad, grad)   (:58)
  v = __fadd_rn(__fmul_rn(v, s.beta2), __fmul_rn(__fmul_rn(omb2, gi), gi));
  // exp_avg.mul_(beta1).add_(1 - beta1, grad)                (:59)
  m = __fadd_rn(__fmul_rn(m, s.beta1), __fmul_rn(omb1, gi));
  if (s.mode != 0) {
    if (s.wd_lr != 0.f) p = __fadd_rn(p, __fmul_rn(-s.wd_lr, p));                 // :83 / :89
    if (s.mode == 1) {
      const float denom = __fadd_rn(__fsqrt_rn(v), s.eps);                        // :84
      p = __fadd_rn(p, __fmul_rn(-s.step_lr, __fdiv_rn(m, denom)));               // :85 addcdiv
    } else {
      p = __fadd_rn(p, __fmul_rn(-s.step_lr, m));                                 // :90
    }
  }
}

// One pass over the span: reads p, g, m, v, writes p, m, v (and g = 0 when asked): 28 (32) bytes per parameter,
// as 16-byte vectors when the four pointers allow it.
__device__ __forceinline__ void radam_span(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                           float* __restrict__ v, int64_t n, const RadamScalars& s, bool vec) {
  const float omb1 = 1.f - s.beta1, omb2 = 1.f - s.beta2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = vec ? (n >> 2) : 0;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = tid; i < n4; i += nthreads) {
    const float4 gg = g4[i];
    float4 mm = m4[i], vv = v4[i], pp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s.mode != 0) pp = p4[i];
    radam_one(pp.x, gg.x, mm.x, vv.x, s, omb1, omb2);
    radam_one(pp.y, gg.y, mm.y, vv.y, s, omb1, omb2);
    radam_one(pp.z, gg.z, mm.z, vv.z, s, omb1, omb2);
    radam_one(pp.w, gg.w, mm.w, vv.w, s, omb1, omb2);
    m4[i] = mm;
    v4[i] = vv;
    if (s.mode != 0) p4[i] = pp;
    if (s.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nthreads) {
    float pi = (s.mode != 0) ? p[i] : 0.f, mi = m[i], vi = v[i];
    radam_one(pi, g[i], mi, vi, s, omb1, omb2);
    m[i] = mi;
    v[i] = vi;
    if (s.mode != 0) p[i] = pi;
    if (s.zero_grad) g[i] = 0.f;
  }
}

__global__ void __launch_bounds__(256)
radam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
             RadamScalars s, int vec) {
  radam_span(p, g, m, v, n, s, vec != 0);
}

// Same update with the step-dependent scalars read from device memory, so that the launch can live inside
// a CUDA graph: hp = {beta1, beta2, eps, wd*lr, step_size*lr, grad_scale, mode, zero_grad}.
__global__ void __launch_bounds__(256)
radam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 const float* __restrict__ hp, int vec) {
  RadamScalars s;
  s.beta1 = __ldg(hp);
  s.beta2 = __ldg(hp + 1);
  s.eps = __ldg(hp + 2);
  s.wd_lr = __ldg(hp + 3);
  s.step_lr = __ldg(hp + 4);
  s.grad_scale = __ldg(hp + 5);
  s.mode = (int)__ldg(hp + 6);
  s.zero_grad = (int)__ldg(hp + 7);
  radam_span(p, g, m, v, n, s, vec != 0);
}

static inline int all_aligned16(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15u) == 0;
}

static inline unsigned radam_grid(int64_t n, int vec) {
  const int64_t work = vec ? (n + 3) / 4 : n;
  const int64_t want = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace hn

extern "C" int hn_radam_step_dev(float* p, float* g, float* m, float* v, int64_t n, const float* hp, void* stream) {
  HN_REQUIRE(n >= 0, "hn_radam_step_dev: negative n");
  if (n == 0) return 0;
  HN_REQUIRE(p && g && m && v && hp, "hn_radam_step_dev: null pointer");
  const int vec = hn::all_aligned16(p, g, m, v);
  hn::radam_dev_kernel<<<hn::radam_grid(n, vec), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hp, vec);
  return hn::check_launch("radam_dev_kernel");
}

extern "C" int hn_radam_step(float* p, float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps,
                             double lr, double weight_decay, double step_size, int mode, float grad_scale,
                             int zero_grad, void* stream) {
  HN_REQUIRE(n >= 0, "hn_radam_step: negative n");
  HN_REQUIRE(mode >= 0 && mode <= 2, "hn_radam_step: mode must be 0, 1 or 2");
  if (n == 0) return 0;
  HN_REQUIRE(p && g && m && v, "hn_radam_step: null pointer");
  hn::RadamScalars s;
  s.beta1 = beta1;
  s.beta2 = beta2;
  s.eps = eps;
  // the reference forms the products in double (python floats) and hands them to ATen as scalars (:83, :85)
  s.wd_lr = (float)(weight_decay * lr);
  s.step_lr = (float)(step_size * lr);
  s.grad_scale = grad_scale;
  s.mode = mode;
  s.zero_grad = zero_grad ? 1 : 0;
  const int vec = hn::all_aligned16(p, g, m, v);
  hn::radam_kernel<<<hn::radam_grid(n, vec), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, s, vec);
  return hn::check_launch("radam_kernel");
}
