// mse_loss.cu -- img2mse (reference run_nerf_helpers.py:24: mean((x - y) ** 2)) as one launch each way.
//
// The photometric loss of a 1024-ray batch is 3072 numbers; as ATen elementwise ops it is three launches forward
// (sub, pow, mean) and four backward, twice per step (coarse and fine image) -- more launch than work in a
// 0.6 ms training step.  One CTA does the whole reduction in a fixed order, so the value is the same on every call.
#include "common.cuh"

namespace hn {

constexpr int kMseThreads = 1024;

__global__ void __launch_bounds__(kMseThreads)
mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ out) {
  __shared__ float warp_sum[kMseThreads / 32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += kMseThreads) {
    const float d = __ldg(a + i) - __ldg(b + i);
    acc = fmaf(d, d, acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = warp_sum[threadIdx.x];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (threadIdx.x == 0) out[0] = s / (float)n;
  }
}

// da = gout * 2 (a - b) / n; db = -da (either may be NULL)
__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, const float* __restrict__ gout,
               float* __restrict__ da, float* __restrict__ db) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = __ldg(gout) * 2.f / (float)n;
  const float v = g * (__ldg(a + i) - __ldg(b + i));
  if (da) da[i] = v;
  if (db) db[i] = -v;
}

}  // namespace hn

extern "C" {

int hn_mse_fwd(const float* a, const float* b, int64_t n, float* out, void* stream) {
  HN_REQUIRE(n >= 1, "hn_mse_fwd: n must be >= 1");
  HN_REQUIRE(a && b && out, "hn_mse_fwd: null pointer");
  hn::mse_fwd_kernel<<<1, hn::kMseThreads, 0, (cudaStream_t)stream>>>(a, b, n, out);
  return hn::check_launch("mse_fwd_kernel");
}

int hn_mse_bwd(const float* a, const float* b, int64_t n, const float* gout, float* da, float* db, void* stream) {
  HN_REQUIRE(n >= 1, "hn_mse_bwd: n must be >= 1");
  HN_REQUIRE(a && b && gout && (da || db), "hn_mse_bwd: null pointer");
  const int64_t grid = (n + 255) / 256;
  HN_REQUIRE(grid <= 0x7fffffffll, "hn_mse_bwd: n exceeds the 1-D grid limit");
  hn::mse_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(a, b, n, gout, da, db);
  return hn::check_launch("mse_bwd_kernel");
}

}  // extern "C"
