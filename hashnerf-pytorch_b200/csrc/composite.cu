// composite.cu -- volume compositing (raw2outputs) forward and backward, one warp per ray.
//
// Replaces raw2outputs (reference run_nerf_helpers.py:577-628): distances, sigmoid colours,
// alpha = 1 - exp(-relu(sigma + noise) * dist * |d|), exclusive cumulative product of (1 - alpha + 1e-10),
// weights, rgb / depth / disparity / accumulation maps, optional white background, and the entropy of
// Categorical(probs = [w_0 .. w_{S-1}, 1 - sum(w) + 1e-6]) ("sparsity loss", :621-626).
//
// The transmittance scan runs as a 32-wide shuffle scan per chunk of 32 samples with the running product
// carried across chunks, so S is arbitrary (64 coarse, 192 fine in the chair config).
#include "common.cuh"

namespace hn {

constexpr unsigned kFull = 0xffffffffu;
constexpr float kEps32 = 1.1920928955078125e-07f;  // torch.finfo(float32).eps, used by clamp_probs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

// exclusive prefix product over the warp; `total` receives the product of all 32 lanes
__device__ __forceinline__ float warp_excl_prod(float v, int lane, float& total) {
  float incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float u = __shfl_up_sync(kFull, incl, off);
    if (lane >= off) incl *= u;
  }
  total = __shfl_sync(kFull, incl, 31);
  const float ex = __shfl_up_sync(kFull, incl, 1);
  return lane == 0 ? 1.f : ex;
}

// exclusive SUFFIX sum over the warp (sum of lanes > lane); `total` = sum of all lanes
__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane, float& total) {
  float incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float u = __shfl_down_sync(kFull, incl, off);
    if (lane + off < 32) incl += u;
  }
  total = __shfl_sync(kFull, incl, 0);
  const float ex = __shfl_down_sync(kFull, incl, 1);
  return lane == 31 ? 0.f : ex;
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

struct SampleTerms {
  float alpha, e, t, dist;
  bool positive;  // relu gate of sigma + noise
};

__device__ __forceinline__ SampleTerms sample_terms(float sigma_plus_noise, float z, float z_next, bool last,
                                                    float norm) {
  SampleTerms s;
  const float gap = last ? 1e10f : (z_next - z);             // :592-593
  s.dist = gap * norm;                                       // :595
  s.positive = sigma_plus_noise > 0.f;
  const float act = fmaxf(sigma_plus_noise, 0.f);            // relu
  s.e = expf(-act * s.dist);
  s.alpha = 1.f - s.e;                                       // :590
  s.t = __fadd_rn(__fsub_rn(1.f, s.alpha), 1e-10f);          // 1 - alpha + 1e-10 (:611)
  return s;
}

// d/dp of -p*log(clamp(p, eps, 1-eps)) as autograd evaluates it (clamp passes gradient inside its range)
__device__ __forceinline__ float entropy_term_grad(float p) {
  const float pc = fminf(fmaxf(p, kEps32), 1.f - kEps32);
  const bool inside = (p >= kEps32) && (p <= 1.f - kEps32);
  return -logf(pc) - (inside ? p / pc : 0.f);
}

__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     const float* __restrict__ noise, int64_t R, int S, int white_bkgd, float* __restrict__ rgb_out,
                     float* __restrict__ disp_out, float* __restrict__ acc_out, float* __restrict__ weights,
                     float* __restrict__ depth_out, float* __restrict__ entropy_out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const float dx = __ldg(rays_d + r * 3), dy = __ldg(rays_d + r * 3 + 1), dz = __ldg(rays_d + r * 3 + 2);
  const float norm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float* zr = z + r * S;
  const float4* rawr = reinterpret_cast<const float4*>(raw) + r * S;
  float* wr = weights + r * S;

  float T_run = 1.f, s_r = 0.f, s_g = 0.f, s_b = 0.f, s_w = 0.f, s_wz = 0.f;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const bool valid = s < S;
    float t = 1.f, alpha = 0.f, zs = 0.f;
    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      zs = __ldg(zr + s);
      const float zn = (s + 1 < S) ? __ldg(zr + s + 1) : 0.f;
      rv = __ldg(rawr + s);
      const float sg = rv.w + (noise ? __ldg(noise + r * S + s) : 0.f);
      const SampleTerms st = sample_terms(sg, zs, zn, s == S - 1, norm);
      t = st.t;
      alpha = st.alpha;
    }
    float total;
    const float ex = warp_excl_prod(t, lane, total);
    const float w = alpha * (T_run * ex);
    T_run *= total;
    if (valid) {
      wr[s] = w;
      s_r = fmaf(w, sigmoidf_ref(rv.x), s_r);
      s_g = fmaf(w, sigmoidf_ref(rv.y), s_g);
      s_b = fmaf(w, sigmoidf_ref(rv.z), s_b);
      s_w += w;
      s_wz = fmaf(w, zs, s_wz);
    }
  }
  s_r = warp_sum(s_r);
  s_g = warp_sum(s_g);
  s_b = warp_sum(s_b);
  const float acc = warp_sum(s_w);
  const float depth = warp_sum(s_wz) / acc;  // :614 (NaN when acc == 0, as the reference)

  // entropy of [w, 1 - acc + 1e-6] normalised by its sum (:623)
  const float q_last = (1.f - acc) + 1e-6f;
  const float Z = acc + q_last;
  float h = 0.f;
  __syncwarp();
  for (int s = lane; s < S; s += 32) {
    const float p = wr[s] / Z;
    h = fmaf(p, logf(fminf(fmaxf(p, kEps32), 1.f - kEps32)), h);
  }
  h = warp_sum(h);
  if (lane == 0) {
    const float p = q_last / Z;
    h = fmaf(p, logf(fminf(fmaxf(p, kEps32), 1.f - kEps32)), h);
    const float bg = white_bkgd ? (1.f - acc) : 0.f;  // :618-619
    rgb_out[r * 3] = s_r + bg;
    rgb_out[r * 3 + 1] = s_g + bg;
    rgb_out[r * 3 + 2] = s_b + bg;
    acc_out[r] = acc;
    depth_out[r] = depth;
    disp_out[r] = (depth != depth) ? depth : 1.f / fmaxf(1e-10f, depth);  // :615, NaN-propagating max
    entropy_out[r] = -h;
  }
}

// Backward.  d_raw doubles as scratch: pass 1 stores (w, T, e, dist) per sample in it, pass 2 reads them
// back (same lane, same address) and overwrites with the gradient.
__global__ void __launch_bounds__(256)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     const float* __restrict__ noise, int64_t R, int S, int white_bkgd,
                     const float* __restrict__ d_rgb, const float* __restrict__ d_disp,
                     const float* __restrict__ d_acc, const float* __restrict__ d_weights,
                     const float* __restrict__ d_depth, const float* __restrict__ d_entropy,
                     float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const float dx = __ldg(rays_d + r * 3), dy = __ldg(rays_d + r * 3 + 1), dz = __ldg(rays_d + r * 3 + 2);
  const float norm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float* zr = z + r * S;
  const float4* rawr = reinterpret_cast<const float4*>(raw) + r * S;
  float4* gr = reinterpret_cast<float4*>(d_raw) + r * S;

  // ---- pass 1: recompute the forward, stash per-sample terms
  float T_run = 1.f, s_w = 0.f, s_wz = 0.f;
  for (int base = 0; base < S; base += 32) {
    const int s = base + lane;
    const bool valid = s < S;
    float t = 1.f, alpha = 0.f, zs = 0.f, e = 1.f, dist = 0.f;
    if (valid) {
      zs = __ldg(zr + s);
      const float zn = (s + 1 < S) ? __ldg(zr + s + 1) : 0.f;
      const float sg = __ldg(raw + (r * S + s) * 4 + 3) + (noise ? __ldg(noise + r * S + s) : 0.f);
      const SampleTerms st = sample_terms(sg, zs, zn, s == S - 1, norm);
      t = st.t;
      alpha = st.alpha;
      e = st.e;
      dist = st.dist;
    }
    float total;
    const float ex = warp_excl_prod(t, lane, total);
    const float T = T_run * ex;
    const float w = alpha * T;
    T_run *= total;
    if (valid) {
      gr[s] = make_float4(w, T, e, dist);
      s_w += w;
      s_wz = fmaf(w, zs, s_wz);
    }
  }
  const float acc = warp_sum(s_w);
  const float depth = warp_sum(s_wz) / acc;
  const float q_last = (1.f - acc) + 1e-6f;
  const float Z = acc + q_last;

  const float g_r = d_rgb ? __ldg(d_rgb + r * 3) : 0.f;
  const float g_g = d_rgb ? __ldg(d_rgb + r * 3 + 1) : 0.f;
  const float g_b = d_rgb ? __ldg(d_rgb + r * 3 + 2) : 0.f;
  const float g_acc = d_acc ? __ldg(d_acc + r) : 0.f;
  const float g_ent = d_entropy ? __ldg(d_entropy + r) : 0.f;
  float g_depth = d_depth ? __ldg(d_depth + r) : 0.f;
  if (d_disp) {
    const float gd = __ldg(d_disp + r);
    // disp = 1 / max(1e-10, depth): the gradient reaches depth only where depth is the larger operand
    if (depth > 1e-10f) g_depth -= gd / (depth * depth);
    else if (depth != depth) g_depth += gd * depth;  // keep NaN poisoning identical to autograd
  }
  const float bg = white_bkgd ? (g_r + g_g + g_b) : 0.f;  // d(1 - acc)/dw = -1 on each channel
  const float h_last = entropy_term_grad(q_last / Z);
  const bool use_depth = (d_depth != nullptr) || (d_disp != nullptr);
  const bool use_ent = d_entropy != nullptr;
  __syncwarp();

  // ---- pass 2: reverse sweep with the running suffix sum  sum_{j>s} G_j w_j
  float suffix = 0.f;
  const int n_chunks = (S + 31) / 32;
  for (int c = n_chunks - 1; c >= 0; --c) {
    const int s = c * 32 + lane;
    const bool valid = s < S;
    float G = 0.f, w = 0.f, T = 0.f, e = 1.f, dist = 0.f;
    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
    float cr = 0.f, cg = 0.f, cb = 0.f;
    bool positive = false;
    if (valid) {
      const float4 st = gr[s];
      w = st.x;
      T = st.y;
      e = st.z;
      dist = st.w;
      rv = __ldg(rawr + s);
      positive = (rv.w + (noise ? __ldg(noise + r * S + s) : 0.f)) > 0.f;
      cr = sigmoidf_ref(rv.x);
      cg = sigmoidf_ref(rv.y);
      cb = sigmoidf_ref(rv.z);
      G = g_r * cr + g_g * cg + g_b * cb - bg + g_acc;
      if (use_depth) G += g_depth * ((__ldg(zr + s) - depth) / acc);
      if (use_ent) G += g_ent * ((entropy_term_grad(w / Z) - h_last) / Z);
      if (d_weights) G += __ldg(d_weights + r * S + s);
    }
    float total;
    const float ex = warp_excl_suffix_sum(G * w, lane, total);
    const float after = suffix + ex;  // sum over samples strictly behind this one
    suffix += total;
    if (valid) {
      const float alpha = 1.f - e;
      const float t = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
      const float d_alpha = G * T - after / t;
      const float d_sigma = positive ? d_alpha * (dist * e) : 0.f;
      gr[s] = make_float4(g_r * w * cr * (1.f - cr), g_g * w * cg * (1.f - cg), g_b * w * cb * (1.f - cb), d_sigma);
    }
  }
}

}  // namespace hn

extern "C" {

int hn_composite_fwd(const float* raw, const float* z, const float* rays_d, const float* noise, int64_t R, int S,
                     int white_bkgd, float* rgb, float* disp, float* acc, float* weights, float* depth, float* entropy,
                     void* stream) {
  HN_REQUIRE(R >= 0 && S >= 1, "hn_composite_fwd: bad shape");
  if (R == 0) return 0;
  HN_REQUIRE(raw && z && rays_d && rgb && disp && acc && weights && depth && entropy, "hn_composite_fwd: null pointer");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15u) == 0, "hn_composite_fwd: raw must be 16-byte aligned");
  const int64_t threads = R * 32;
  hn::composite_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      raw, z, rays_d, noise, R, S, white_bkgd, rgb, disp, acc, weights, depth, entropy);
  return hn::check_launch("composite_fwd_kernel");
}

int hn_composite_bwd(const float* raw, const float* z, const float* rays_d, const float* noise, int64_t R, int S,
                     int white_bkgd, const float* d_rgb, const float* d_disp, const float* d_acc,
                     const float* d_weights, const float* d_depth, const float* d_entropy, float* d_raw,
                     void* stream) {
  HN_REQUIRE(R >= 0 && S >= 1, "hn_composite_bwd: bad shape");
  if (R == 0) return 0;
  HN_REQUIRE(raw && z && rays_d && d_raw, "hn_composite_bwd: null pointer");
  HN_REQUIRE(((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(d_raw)) & 15u) == 0,
             "hn_composite_bwd: raw and d_raw must be 16-byte aligned");
  const int64_t threads = R * 32;
  hn::composite_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      raw, z, rays_d, noise, R, S, white_bkgd, d_rgb, d_disp, d_acc, d_weights, d_depth, d_entropy, d_raw);
  return hn::check_launch("composite_bwd_kernel");
}

}  // extern "C"
