// tv_loss.cu -- total-variation regulariser of one hash level ("next" row 2).
//
// Replaces total_variation_loss (reference loss.py:11-43): features are gathered at the (cube+1)^3 hashed
// vertices of a random cube of the level's grid and the squared forward differences along x, y, z are summed and
// divided by the cube size.  The reference materialises the index cube, the gathered cube and three difference
// tensors (~15 launches per level, 16 levels per step); here the forward is one launch and the backward one
// launch that scatters straight into the level's gradient slab.
#include "common.cuh"

namespace hn {

__device__ __forceinline__ uint32_t tv_hash(int64_t x, int64_t y, int64_t z, uint32_t mask) {
  return hash3((uint32_t)x, (uint32_t)y, (uint32_t)z, mask);
}

// blockIdx.y = level: `cubes` == nullptr -> one level with cube size `cube`; otherwise level l uses cubes[l],
// origin + 3*l, the table slab l and out[l] (all levels of a sweep in one launch).
template <int F>
__global__ void __launch_bounds__(256)
tv_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ origin, int cube,
              const int32_t* __restrict__ cubes, int log2T, float* __restrict__ out) {
  if (cubes != nullptr) {
    const int l = blockIdx.y;
    cube = __ldg(cubes + l);
    origin += 3 * l;
    table += ((size_t)l << log2T) * F;
    out += l;
  }
  const int n1 = cube + 1;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if ((int64_t)blockIdx.x * blockDim.x >= (int64_t)n1 * n1 * n1) return;  // CTA beyond this level's cube (uniform)
  const uint32_t mask = (1u << log2T) - 1u;
  float acc = 0.f;
  if (v < n1 * n1 * n1) {
    const int k = v % n1, j = (v / n1) % n1, i = v / (n1 * n1);  // k fastest: the reference's meshgrid 'ij' order
    const int64_t x = origin[0] + i, y = origin[1] + j, z = origin[2] + k;
    float e[F], d[F];
    const uint32_t h = tv_hash(x, y, z, mask);
#pragma unroll
    for (int f = 0; f < F; ++f) e[f] = __ldg(table + (size_t)h * F + f);
    if (i < cube) {  // loss.py:39
      const uint32_t hn = tv_hash(x + 1, y, z, mask);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        d[f] = __ldg(table + (size_t)hn * F + f) - e[f];
        acc = fmaf(d[f], d[f], acc);
      }
    }
    if (j < cube) {  // :40
      const uint32_t hn = tv_hash(x, y + 1, z, mask);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        d[f] = __ldg(table + (size_t)hn * F + f) - e[f];
        acc = fmaf(d[f], d[f], acc);
      }
    }
    if (k < cube) {  // :41
      const uint32_t hn = tv_hash(x, y, z + 1, mask);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        d[f] = __ldg(table + (size_t)hn * F + f) - e[f];
        acc = fmaf(d[f], d[f], acc);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ float warp_part[8];
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += warp_part[w];
    atomicAdd(out, s / (float)cube);  // :43
  }
}

// d tv / d e(v) = (2 / cube) * sum over the up-to-6 grid neighbours n inside the cube of (e(v) - e(n))
template <int F>
__global__ void __launch_bounds__(256)
tv_bwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ origin, int cube,
              const int32_t* __restrict__ cubes, int log2T, const float* __restrict__ gout,
              float* __restrict__ dtable) {
  if (cubes != nullptr) {
    const int l = blockIdx.y;
    cube = __ldg(cubes + l);
    origin += 3 * l;
    table += ((size_t)l << log2T) * F;
    dtable += ((size_t)l << log2T) * F;
    gout += l;
  }
  const int n1 = cube + 1;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n1 * n1 * n1) return;
  const uint32_t mask = (1u << log2T) - 1u;
  const int k = v % n1, j = (v / n1) % n1, i = v / (n1 * n1);
  const int64_t x = origin[0] + i, y = origin[1] + j, z = origin[2] + k;
  const uint32_t h = tv_hash(x, y, z, mask);
  float e[F], g[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    e[f] = __ldg(table + (size_t)h * F + f);
    g[f] = 0.f;
  }
  auto edge = [&](int64_t nx, int64_t ny, int64_t nz) {
    const uint32_t hn = tv_hash(nx, ny, nz, mask);
#pragma unroll
    for (int f = 0; f < F; ++f) g[f] += e[f] - __ldg(table + (size_t)hn * F + f);
  };
  if (i < cube) edge(x + 1, y, z);
  if (i > 0) edge(x - 1, y, z);
  if (j < cube) edge(x, y + 1, z);
  if (j > 0) edge(x, y - 1, z);
  if (k < cube) edge(x, y, z + 1);
  if (k > 0) edge(x, y, z - 1);
  const float scale = 2.f * __ldg(gout) / (float)cube;
#pragma unroll
  for (int f = 0; f < F; ++f) atomicAdd(dtable + (size_t)h * F + f, scale * g[f]);
}

}  // namespace hn

extern "C" {

int hn_tv_loss_fwd(const float* table, const int64_t* origin, int cube, int log2T, int F, float* out, void* stream) {
  HN_REQUIRE(cube >= 1 && cube <= 255, "hn_tv_loss_fwd: cube size must be in [1,255]");
  HN_REQUIRE(log2T >= 1 && log2T <= 30, "hn_tv_loss_fwd: log2T out of range");
  HN_REQUIRE(F == 1 || F == 2 || F == 4, "hn_tv_loss_fwd: F must be 1, 2 or 4");
  HN_REQUIRE(table && origin && out, "hn_tv_loss_fwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), s);
  if (e != cudaSuccess) return hn::fail((int)e, "cudaMemsetAsync(tv out)");
  const int n = (cube + 1) * (cube + 1) * (cube + 1);
  const unsigned grid = (unsigned)((n + 255) / 256);
  switch (F) {
    case 1: hn::tv_fwd_kernel<1><<<grid, 256, 0, s>>>(table, origin, cube, nullptr, log2T, out); break;
    case 2: hn::tv_fwd_kernel<2><<<grid, 256, 0, s>>>(table, origin, cube, nullptr, log2T, out); break;
    default: hn::tv_fwd_kernel<4><<<grid, 256, 0, s>>>(table, origin, cube, nullptr, log2T, out); break;
  }
  return hn::check_launch("tv_fwd_kernel");
}

int hn_tv_loss_bwd(const float* table, const int64_t* origin, int cube, int log2T, int F, const float* gout,
                   float* dtable, void* stream) {
  HN_REQUIRE(cube >= 1 && cube <= 255, "hn_tv_loss_bwd: cube size must be in [1,255]");
  HN_REQUIRE(log2T >= 1 && log2T <= 30, "hn_tv_loss_bwd: log2T out of range");
  HN_REQUIRE(F == 1 || F == 2 || F == 4, "hn_tv_loss_bwd: F must be 1, 2 or 4");
  HN_REQUIRE(table && origin && gout && dtable, "hn_tv_loss_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (cube + 1) * (cube + 1) * (cube + 1);
  const unsigned grid = (unsigned)((n + 255) / 256);
  switch (F) {
    case 1: hn::tv_bwd_kernel<1><<<grid, 256, 0, s>>>(table, origin, cube, nullptr, log2T, gout, dtable); break;
    case 2: hn::tv_bwd_kernel<2><<<grid, 256, 0, s>>>(table, origin, cube, nullptr, log2T, gout, dtable); break;
    default: hn::tv_bwd_kernel<4><<<grid, 256, 0, s>>>(table, origin, cube, nullptr, log2T, gout, dtable); break;
  }
  return hn::check_launch("tv_bwd_kernel");
}

int hn_tv_loss_fwd_levels(const float* tables, const int64_t* origins, const int32_t* cubes, int L, int max_cube,
                          int log2T, int F, float* out, void* stream) {
  HN_REQUIRE(L >= 1 && L <= HN_MAX_LEVELS, "hn_tv_loss_fwd_levels: L out of range");
  HN_REQUIRE(max_cube >= 1 && max_cube <= 255, "hn_tv_loss_fwd_levels: cube size must be in [1,255]");
  HN_REQUIRE(log2T >= 1 && log2T <= 30, "hn_tv_loss_fwd_levels: log2T out of range");
  HN_REQUIRE(F == 1 || F == 2 || F == 4, "hn_tv_loss_fwd_levels: F must be 1, 2 or 4");
  HN_REQUIRE(tables && origins && cubes && out, "hn_tv_loss_fwd_levels: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)L, s);
  if (e != cudaSuccess) return hn::fail((int)e, "cudaMemsetAsync(tv out)");
  const int n = (max_cube + 1) * (max_cube + 1) * (max_cube + 1);
  const dim3 grid((unsigned)((n + 255) / 256), (unsigned)L);
  switch (F) {
    case 1: hn::tv_fwd_kernel<1><<<grid, 256, 0, s>>>(tables, origins, 0, cubes, log2T, out); break;
    case 2: hn::tv_fwd_kernel<2><<<grid, 256, 0, s>>>(tables, origins, 0, cubes, log2T, out); break;
    default: hn::tv_fwd_kernel<4><<<grid, 256, 0, s>>>(tables, origins, 0, cubes, log2T, out); break;
  }
  return hn::check_launch("tv_fwd_kernel(levels)");
}

int hn_tv_loss_bwd_levels(const float* tables, const int64_t* origins, const int32_t* cubes, int L, int max_cube,
                          int log2T, int F, const float* gout, float* dtables, void* stream) {
  HN_REQUIRE(L >= 1 && L <= HN_MAX_LEVELS, "hn_tv_loss_bwd_levels: L out of range");
  HN_REQUIRE(max_cube >= 1 && max_cube <= 255, "hn_tv_loss_bwd_levels: cube size must be in [1,255]");
  HN_REQUIRE(log2T >= 1 && log2T <= 30, "hn_tv_loss_bwd_levels: log2T out of range");
  HN_REQUIRE(F == 1 || F == 2 || F == 4, "hn_tv_loss_bwd_levels: F must be 1, 2 or 4");
  HN_REQUIRE(tables && origins && cubes && gout && dtables, "hn_tv_loss_bwd_levels: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (max_cube + 1) * (max_cube + 1) * (max_cube + 1);
  const dim3 grid((unsigned)((n + 255) / 256), (unsigned)L);
  switch (F) {
    case 1: hn::tv_bwd_kernel<1><<<grid, 256, 0, s>>>(tables, origins, 0, cubes, log2T, gout, dtables); break;
    case 2: hn::tv_bwd_kernel<2><<<grid, 256, 0, s>>>(tables, origins, 0, cubes, log2T, gout, dtables); break;
    default: hn::tv_bwd_kernel<4><<<grid, 256, 0, s>>>(tables, origins, 0, cubes, log2T, gout, dtables); break;
  }
  return hn::check_launch("tv_bwd_kernel(levels)");
}

}  // extern "C"
