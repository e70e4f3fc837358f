// dp_exchange.cu -- the data-parallel exchange step as ONE pass over peer memory.
//
// The reference has no distributed code; the north star's exchange is "hash-table and MLP gradients averaged with
// an all-reduce".  Done with a library all-reduce that is three passes over the 64 MiB table gradient per rank and
// step: the collective (read + write), the optimizer (read g, p, m, v; write p, m, v) and the zero-fill of the
// gradient for the next step.  Here every rank owns 1/world of each flat buffer and does, in one kernel:
//
//   g  = sum over ranks of their gradient slice   multimem.ld_reduce.add.v4.f32 through the NVSwitch (NVLS), or
//                                                 one peer load per rank over NVLink when there is no multicast
//   (p, m, v) <- RAdam(p, g / world, m, v)        the owner's slice only: moments are sharded, 1/world of the
//                                                 optimizer traffic per rank
//   p  -> every rank                              multimem.st.v4.f32 (or one peer store per rank)
//   g  = 0 on every rank                          multimem.st of zeros: the zero_grad of the next step
//
// The buffers are symmetric allocations (same size and layout on every rank) whose peer pointers / multicast
// pointers the caller obtains from its allocator (torch.distributed._symmetric_memory in hn_b200/dp.py); this file
// only sees raw pointers.  Ranks meet before (all gradients written) and after (all parameters landed) through
// hn_dp_barrier: one flag per (peer, rank) in a symmetric signal buffer, monotonically increasing epochs, release
// stores / acquire loads at system scope.  A rank that waits longer than ~2 s traps instead of hanging the GPU.
// One rank per GPU only (B200_PROFILING.md: kernels of different ranks that wait on one another must never share
// a GPU).
#include "common.cuh"

namespace hn {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// signal[peer][slot * 64 + rank] = epoch, then wait until every peer has written its flag for this rank
__global__ void __launch_bounds__(64)
dp_barrier_kernel(uint32_t* const* __restrict__ signal_ptrs, int rank, int world, int slot, uint32_t epoch) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();  // everything this GPU wrote before the barrier is visible system-wide first
  st_release_sys(signal_ptrs[t] + slot * 64 + rank, epoch);
  const uint32_t* mine = signal_ptrs[rank] + slot * 64 + t;
  const long long t0 = clock64();
  // epochs only grow: ">= epoch" also accepts a peer that is already one exchange ahead
  while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s at 2 GHz: a peer died; fail loudly rather than hang
  }
}

__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

struct DpScalars {
  float beta1, beta2, eps, wd_lr, step_lr, grad_scale;
  int mode;  // 0 moments only, 1 adaptive, 2 SGD-like (optim.cu); -1 = no optimizer: plain summed all-reduce
};

__device__ __forceinline__ void radam_elem(float& p, float g, float& m, float& v, const DpScalars& s, float omb1,
                                           float omb2) {
  const float gi = g * s.grad_scale;
  v = __fadd_rn(__fmul_rn(v, s.beta2), __fmul_rn(__fmul_rn(omb2, gi), gi));   // radam.py:58
  m = __fadd_rn(__fmul_rn(m, s.beta1), __fmul_rn(omb1, gi));                  // :59
  if (s.mode != 0) {
    if (s.wd_lr != 0.f) p = __fadd_rn(p, __fmul_rn(-s.wd_lr, p));             // :83 / :89
    if (s.mode == 1) p = __fadd_rn(p, __fmul_rn(-s.step_lr, __fdiv_rn(m, __fadd_rn(__fsqrt_rn(v), s.eps))));  // :84-85
    else p = __fadd_rn(p, __fmul_rn(-s.step_lr, m));                          // :90
  }
}

// [begin, end): this rank's slice of the span, in floats, both multiples of 4.  grads / params: device arrays of
// the `world` peer pointers to the span's start on every rank; *_mc: multicast pointers to the same (MC only).
// m, v: this rank's moment buffers for the span (indexed like the span; only the slice is touched).
template <bool MC>
__global__ void __launch_bounds__(256)
dp_reduce_update_kernel(float* const* __restrict__ grads, float* __restrict__ grads_mc, float* const* __restrict__ params,
                        float* __restrict__ params_mc, float* __restrict__ m, float* __restrict__ v, int rank, int world,
                        int64_t begin, int64_t end, const float* __restrict__ hp) {
  DpScalars s;
  s.beta1 = __ldg(hp);
  s.beta2 = __ldg(hp + 1);
  s.eps = __ldg(hp + 2);
  s.wd_lr = __ldg(hp + 3);
  s.step_lr = __ldg(hp + 4);
  s.grad_scale = __ldg(hp + 5);
  s.mode = (int)__ldg(hp + 6);
  const float omb1 = 1.f - s.beta1, omb2 = 1.f - s.beta2;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t n4 = (end - begin) >> 2;
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = begin + 4 * i4;
    float4 g;
    if (MC) {
      g = mc_ld_reduce_add(grads_mc + i);
    } else {
      g = zero;
      for (int r = 0; r < world; ++r) {
        const float4 t = *reinterpret_cast<const float4*>(grads[(rank + r) % world] + i);  // start at home: spreads links
        g.x += t.x;
        g.y += t.y;
        g.z += t.z;
        g.w += t.w;
      }
    }
    float4 out;
    if (s.mode < 0) {
      out = g;  // plain all-reduce: the sum goes back into the gradient buffers
      if (MC) mc_st(grads_mc + i, out);
      else
        for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(grads[(rank + r) % world] + i) = out;
      continue;
    }
    float4 p = *reinterpret_cast<const float4*>(params[rank] + i);
    float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    radam_elem(p.x, g.x, mm.x, vv.x, s, omb1, omb2);
    radam_elem(p.y, g.y, mm.y, vv.y, s, omb1, omb2);
    radam_elem(p.z, g.z, mm.z, vv.z, s, omb1, omb2);
    radam_elem(p.w, g.w, mm.w, vv.w, s, omb1, omb2);
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (MC) {
      if (s.mode != 0) mc_st(params_mc + i, p);
      mc_st(grads_mc + i, zero);
    } else {
      for (int r = 0; r < world; ++r) {
        const int q = (rank + r) % world;
        if (s.mode != 0) *reinterpret_cast<float4*>(params[q] + i) = p;
        *reinterpret_cast<float4*>(grads[q] + i) = zero;
      }
    }
  }
}

int g_dp_grid_per_sm = 8;  // hn_set_tuning("dp_grid_per_sm"): CTAs of 256 threads per SM for the exchange kernel

}  // namespace hn

extern "C" {

int hn_dp_barrier(void* const* signal_ptrs_dev, int rank, int world, int slot, uint32_t epoch, void* stream) {
  HN_REQUIRE(signal_ptrs_dev != nullptr, "hn_dp_barrier: null pointer");
  HN_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "hn_dp_barrier: bad rank / world (<= 64)");
  HN_REQUIRE(slot >= 0 && slot < 4, "hn_dp_barrier: slot must be in [0,4)");
  hn::dp_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t* const*>(signal_ptrs_dev), rank,
                                                            world, slot, epoch);
  return hn::check_launch("dp_barrier_kernel");
}

int hn_dp_reduce_update(void* const* grad_ptrs_dev, float* grad_mc, void* const* param_ptrs_dev, float* param_mc,
                        float* m, float* v, int rank, int world, int64_t n, const float* hp, void* stream) {
  HN_REQUIRE(n >= 0 && (n & 3) == 0, "hn_dp_reduce_update: span length must be a multiple of 4 floats");
  HN_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "hn_dp_reduce_update: bad rank / world");
  HN_REQUIRE(grad_ptrs_dev && hp, "hn_dp_reduce_update: null pointer");
  if (n == 0) return 0;
  // this rank's slice: ceil(n / world) rounded up to 4 floats
  const int64_t chunk = (((n + world - 1) / world) + 3) & ~(int64_t)3;
  const int64_t begin = (int64_t)rank * chunk < n ? (int64_t)rank * chunk : n;
  const int64_t end = begin + chunk < n ? begin + chunk : n;
  if (end <= begin) return 0;
  const int64_t want = (((end - begin) >> 2) + 255) / 256;
  const int64_t cap = (int64_t)hn::sm_count() * (hn::g_dp_grid_per_sm > 0 ? hn::g_dp_grid_per_sm : 8);
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  float* const* g = reinterpret_cast<float* const*>(grad_ptrs_dev);
  float* const* p = reinterpret_cast<float* const*>(param_ptrs_dev);
  if (grad_mc != nullptr)
    hn::dp_reduce_update_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(g, grad_mc, p, param_mc, m, v, rank, world,
                                                                            begin, end, hp);
  else
    hn::dp_reduce_update_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(g, grad_mc, p, param_mc, m, v, rank, world,
                                                                             begin, end, hp);
  return hn::check_launch("dp_reduce_update_kernel");
}

}  // extern "C"
