// sampling.cu -- ray-marching set-up and hierarchical resampling.
//
//   hn_coarse_z         stratified depths            (reference run_nerf_helpers.py:514-536)
//   hn_ray_points       pts = o + d * z              (:538, :552)
//   hn_sample_pdf       inverse-CDF resampling       (:264-307)
//   hn_sort_concat_rows z = sort(cat(z, z_samples))  (:551)
//
// The elementwise parts follow the reference's fp32 op order exactly (no FMA contraction) so they are
// bit-exact; the CDF is a warp scan, so resampled depths agree to fp32 rounding, not bit for bit.
#include "common.cuh"

namespace hn {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float coarse_depth(float near, float far, float t, int lindisp) {
  const float omt = __fsub_rn(1.f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));  // :516
  const float inv = __fadd_rn(__fmul_rn(__fdiv_rn(1.f, near), omt), __fmul_rn(__fdiv_rn(1.f, far), t));
  return __fdiv_rn(1.f, inv);  // :518
}

__global__ void __launch_bounds__(256)
coarse_z_kernel(const float* __restrict__ near, const float* __restrict__ far, int64_t nf_stride,
                const float* __restrict__ t_vals, const float* __restrict__ t_rand, int64_t R, int S, int lindisp,
                float* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  const int64_t r = i / S;
  const int s = (int)(i % S);
  const float nr = __ldg(near + r * nf_stride), fr = __ldg(far + r * nf_stride);
  const float zc = coarse_depth(nr, fr, __ldg(t_vals + s), lindisp);
  if (t_rand == nullptr) {
    z[i] = zc;
    return;
  }
  // stratified jitter inside [lower, upper] (:524-536)
  float lower = zc, upper = zc;
  if (s > 0) lower = __fmul_rn(0.5f, __fadd_rn(zc, coarse_depth(nr, fr, __ldg(t_vals + s - 1), lindisp)));
  if (s < S - 1) upper = __fmul_rn(0.5f, __fadd_rn(coarse_depth(nr, fr, __ldg(t_vals + s + 1), lindisp), zc));
  z[i] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(t_rand + i)));
}

__global__ void __launch_bounds__(256)
ray_points_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t ray_stride,
                  const float* __restrict__ z, int64_t R, int S, float* __restrict__ pts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  const int64_t r = i / S;
  const float zz = __ldg(z + i);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    pts[i * 3 + c] = __fadd_rn(__ldg(rays_o + r * ray_stride + c), __fmul_rn(__ldg(rays_d + r * ray_stride + c), zz));
}

// One warp per ray; the ray's CDF lives in that warp's slice of dynamic shared memory.
__global__ void __launch_bounds__(128)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, const float* __restrict__ u,
                  const float* __restrict__ u_det, int64_t R, int nb, int Ni, float* __restrict__ samples) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= R) return;
  float* cdf = smem + warp * nb;
  const float* wrow = weights + r * (nb - 1);
  const float* brow = bins + r * nb;

  float part = 0.f;
  for (int i = lane; i < nb - 1; i += 32) part += __ldg(wrow + i) + 1e-5f;  // :266
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(kFullMask, part, off);
  const float total = part;  // :267 denominator

  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;  // :269
  for (int base = 0; base < nb - 1; base += 32) {
    const int i = base + lane;
    float v = (i < nb - 1) ? (__ldg(wrow + i) + 1e-5f) / total : 0.f;  // pdf
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float t = __shfl_up_sync(kFullMask, v, off);
      if (lane >= off) v += t;
    }
    if (i < nb - 1) cdf[i + 1] = carry + v;  // :268
    carry += __shfl_sync(kFullMask, v, 31);
  }
  __syncwarp();

  for (int k = lane; k < Ni; k += 32) {
    const float uu = u ? __ldg(u + r * Ni + k) : __ldg(u_det + k);
    // searchsorted(cdf, u, right=True): first index with cdf[idx] > u, in [0, nb]  (:290)
    int lo = 0, hi = nb;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] > uu) hi = mid;
      else lo = mid + 1;
    }
    const int below = max(lo - 1, 0);        // :291
    const int above = min(lo, nb - 1);       // :292
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = __ldg(brow + below), b1 = __ldg(brow + above);
    float denom = __fsub_rn(c1, c0);         // :301
    if (denom < 1e-5f) denom = 1.f;          // :302
    const float t = __fdiv_rn(__fsub_rn(uu, c0), denom);                          // :303
    samples[r * Ni + k] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));         // :304
  }
}

// Bitonic network over a[0..n) (n a power of two) by one warp: every lane owns whole compare-exchange PAIRS
// (pair q <-> i = q with a zero bit inserted at log2 j, partner i | j), so all 32 lanes work in every trip.
__device__ __forceinline__ void bitonic_sort_warp(float* __restrict__ a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 4
      for (int q = lane; q < (n >> 1); q += 32) {
        const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        const int partner = i | j;
        const float x = a[i], y = a[partner];
        const bool up = (i & k) == 0;
        if ((x > y) == up) {
          a[i] = y;
          a[partner] = x;
        }
      }
      __syncwarp();
    }
  }
}

// Fused hierarchical resampling for one ray per warp (run_nerf_helpers.py:547-552, 568):
//   mids = .5 (z[1:] + z[:-1]);  samples = sample_pdf(mids, weights[1:-1], u);  merged = sort(cat(z, samples));
//   z_std = std(samples, unbiased=False).
// Shared memory per warp: cdf[S-1] | mids[S-1] | sort buffer[bufsz = max(npad, S + npad_s)].
__global__ void __launch_bounds__(128)
resample_kernel(const float* __restrict__ z, const float* __restrict__ weights, const float* __restrict__ u,
                const float* __restrict__ u_det, int64_t R, int S, int Ni, int npad, int npad_s, int bufsz,
                float* __restrict__ samples, float* __restrict__ merged, float* __restrict__ z_std) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= R) return;
  const int nb = S - 1;
  float* cdf = smem + (size_t)warp * (2 * nb + bufsz);
  float* mids = cdf + nb;
  float* buf = mids + nb;
  const float* zr = z + r * S;
  const float* wr = weights + r * S + 1;  // weights[..., 1:-1]

  for (int i = lane; i < nb; i += 32) mids[i] = __fmul_rn(0.5f, __fadd_rn(__ldg(zr + i + 1), __ldg(zr + i)));  // :547
  float part = 0.f;
  for (int i = lane; i < nb - 1; i += 32) part += __ldg(wr + i) + 1e-5f;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(kFullMask, part, off);
  const float total = part;
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < nb - 1; base += 32) {
    const int i = base + lane;
    float v = (i < nb - 1) ? (__ldg(wr + i) + 1e-5f) / total : 0.f;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float t = __shfl_up_sync(kFullMask, v, off);
      if (lane >= off) v += t;
    }
    if (i < nb - 1) cdf[i + 1] = carry + v;
    carry += __shfl_sync(kFullMask, v, 31);
  }
  // sort buffer: the S coarse depths, then the Ni new samples, then +inf padding
  for (int i = lane; i < S; i += 32) buf[i] = __ldg(zr + i);
  __syncwarp();

  float s1 = 0.f;
  for (int k = lane; k < Ni; k += 32) {
    const float uu = u ? __ldg(u + r * Ni + k) : __ldg(u_det + k);
    int lo = 0, hi = nb;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] > uu) hi = mid;
      else lo = mid + 1;
    }
    const int below = max(lo - 1, 0), above = min(lo, nb - 1);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = mids[below], b1 = mids[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fsub_rn(uu, c0), denom);
    const float smp = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    samples[r * Ni + k] = smp;
    buf[S + k] = smp;
    s1 += smp;
  }
  if (z_std != nullptr) {  // population standard deviation of the new samples (:568)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s1 += __shfl_xor_sync(kFullMask, s1, off);
    const float mean = s1 / (float)Ni;
    __syncwarp();
    float s2 = 0.f;
    for (int k = lane; k < Ni; k += 32) {
      const float d = buf[S + k] - mean;
      s2 = fmaf(d, d, s2);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(kFullMask, s2, off);
    if (lane == 0) z_std[r] = sqrtf(s2 / (float)Ni);
  }
  __syncwarp();
  // The coarse depths come sorted (stratified sampling) and only the Ni new samples are not: sort those alone
  // (a network over npad_s <= 2 Ni values instead of npad >= S + Ni) and merge by rank -- an element's place in
  // the merged row is its index in its own sequence plus the number of elements of the other sequence before it
  // (strictly smaller for the depths, smaller-or-equal for the samples: a bijection also with ties).  The result
  // is the same sorted row; rows whose depths are not sorted, or that hold a NaN, take the full network.
  bool fast = npad_s > 0;
  for (int i = lane; i < S - 1; i += 32) fast = fast && (buf[i] <= buf[i + 1]);
  for (int k = lane; k < Ni; k += 32) fast = fast && (buf[S + k] == buf[S + k]);
  fast = __all_sync(kFullMask, fast);
  if (fast) {
    float* sb = buf + S;
    for (int i = Ni + lane; i < npad_s; i += 32) sb[i] = __int_as_float(0x7f800000);
    __syncwarp();
    bitonic_sort_warp(sb, npad_s, lane);
    float* out = merged + r * (S + Ni);
    for (int i = lane; i < S; i += 32) {
      const float v = buf[i];
      int lo = 0, hi = Ni;
      while (lo < hi) {  // samples strictly below v
        const int mid = (lo + hi) >> 1;
        if (sb[mid] < v) lo = mid + 1;
        else hi = mid;
      }
      out[i + lo] = v;
    }
    for (int k = lane; k < Ni; k += 32) {
      const float v = sb[k];
      int lo = 0, hi = S;
      while (lo < hi) {  // depths at or below v
        const int mid = (lo + hi) >> 1;
        if (buf[mid] <= v) lo = mid + 1;
        else hi = mid;
      }
      out[k + lo] = v;
    }
    return;
  }
  for (int i = S + Ni + lane; i < npad; i += 32) buf[i] = __int_as_float(0x7f800000);
  __syncwarp();
  bitonic_sort_warp(buf, npad, lane);
  for (int i = lane; i < S + Ni; i += 32) merged[r * (S + Ni) + i] = buf[i];
}

// Block per row, bitonic network over the row padded to a power of two with +inf.
__global__ void sort_concat_rows_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb,
                                        int npad, float* __restrict__ out) {
  extern __shared__ float row[];
  const int64_t r = blockIdx.x;
  const int n = na + nb;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    float v = __int_as_float(0x7f800000);
    if (i < na) v = __ldg(a + r * na + i);
    else if (i < n) v = __ldg(b + r * nb + (i - na));
    row[i] = v;
  }
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int partner = i ^ j;
        if (partner > i) {
          const float x = row[i], y = row[partner];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            row[i] = y;
            row[partner] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[r * n + i] = row[i];
}

}  // namespace hn

extern "C" {

int hn_coarse_z(const float* near, const float* far, int64_t nf_stride, const float* t_vals, const float* t_rand,
                int64_t R, int S, int lindisp, float* z, void* stream) {
  HN_REQUIRE(R >= 0 && S >= 1, "hn_coarse_z: bad shape");
  if (R == 0) return 0;
  HN_REQUIRE(near && far && t_vals && z, "hn_coarse_z: null pointer");
  const int64_t n = R * S;
  hn::coarse_z_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(near, far, nf_stride, t_vals,
                                                                                     t_rand, R, S, lindisp, z);
  return hn::check_launch("coarse_z_kernel");
}

int hn_ray_points(const float* rays_o, const float* rays_d, int64_t ray_stride, const float* z, int64_t R, int S,
                  float* pts, void* stream) {
  HN_REQUIRE(R >= 0 && S >= 1 && ray_stride >= 3, "hn_ray_points: bad shape");
  if (R == 0) return 0;
  HN_REQUIRE(rays_o && rays_d && z && pts, "hn_ray_points: null pointer");
  const int64_t n = R * S;
  hn::ray_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, ray_stride, z,
                                                                                       R, S, pts);
  return hn::check_launch("ray_points_kernel");
}

int hn_sample_pdf(const float* bins, const float* weights, const float* u, const float* u_det, int64_t R, int nb,
                  int Ni, float* samples, void* stream) {
  HN_REQUIRE(R >= 0 && nb >= 2 && Ni >= 1, "hn_sample_pdf: bad shape");
  HN_REQUIRE(nb <= 8192, "hn_sample_pdf: at most 8192 bins per ray");
  if (R == 0) return 0;
  HN_REQUIRE(bins && weights && samples && (u || u_det), "hn_sample_pdf: null pointer");
  const int warps = 4;
  const size_t smem = (size_t)warps * nb * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hn::sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return hn::fail((int)e, "cudaFuncSetAttribute(sample_pdf_kernel)");
  }
  hn::sample_pdf_kernel<<<(unsigned)((R + warps - 1) / warps), warps * 32, smem, (cudaStream_t)stream>>>(
      bins, weights, u, u_det, R, nb, Ni, samples);
  return hn::check_launch("sample_pdf_kernel");
}

int hn_resample(const float* z, const float* weights, const float* u, const float* u_det, int64_t R, int S, int Ni,
                float* samples, float* merged, float* z_std, void* stream) {
  HN_REQUIRE(R >= 0 && S >= 3 && Ni >= 1, "hn_resample: need S >= 3 and Ni >= 1");
  HN_REQUIRE(S + Ni <= 2048, "hn_resample: S + Ni must be <= 2048");
  if (R == 0) return 0;
  HN_REQUIRE(z && weights && samples && merged && (u || u_det), "hn_resample: null pointer");
  int npad = 2;
  while (npad < S + Ni) npad <<= 1;
  int npad_s = 2;
  while (npad_s < Ni) npad_s <<= 1;
  const int bufsz = npad > S + npad_s ? npad : S + npad_s;
  const int warps = 4;
  const size_t smem = (size_t)warps * (2 * (S - 1) + bufsz) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hn::resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return hn::fail((int)e, "cudaFuncSetAttribute(resample_kernel)");
  }
  hn::resample_kernel<<<(unsigned)((R + warps - 1) / warps), warps * 32, smem, (cudaStream_t)stream>>>(
      z, weights, u, u_det, R, S, Ni, npad, npad_s, bufsz, samples, merged, z_std);
  return hn::check_launch("resample_kernel");
}

int hn_sort_concat_rows(const float* a, int na, const float* b, int nb, int64_t R, float* out, void* stream) {
  HN_REQUIRE(R >= 0 && na >= 0 && nb >= 0 && na + nb >= 1, "hn_sort_concat_rows: bad shape");
  HN_REQUIRE(na + nb <= 2048, "hn_sort_concat_rows: rows longer than 2048 are not supported");
  HN_REQUIRE(R <= 0x7fffffff, "hn_sort_concat_rows: too many rows");
  if (R == 0) return 0;
  HN_REQUIRE((a || na == 0) && (b || nb == 0) && out, "hn_sort_concat_rows: null pointer");
  int npad = 2;
  while (npad < na + nb) npad <<= 1;
  const int threads = npad / 2 < 32 ? 32 : (npad / 2 > 1024 ? 1024 : npad / 2);
  hn::sort_concat_rows_kernel<<<(unsigned)R, threads, npad * sizeof(float), (cudaStream_t)stream>>>(a, na, b, nb, npad,
                                                                                                     out);
  return hn::check_launch("sort_concat_rows_kernel");
}

}  // extern "C"
