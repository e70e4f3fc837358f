// mlp.cu -- NeRFSmall sigma/colour MLP, forward and backward, fp32 FFMA "parity" implementation.
//
// Replaces NeRFSmall.forward (reference models.py:151-174) as instantiated at run_nerf_helpers.py:79-84:
//   h1 = relu(enc[32] . W0^T)  (64)      h2 = h1 . W1^T (16)   sigma = h2[0], geo = h2[1:16]
//   c  = [views(16) | geo(15)]           h3 = relu(c . W2^T)   h4 = relu(h3 . W3^T)   rgb = h4 . W4^T
//   out = [rgb(3) | sigma(1)]            (no biases, no output activation)
// and fuses the expand/cat of run_network (run_nerf_helpers.py:219-222): the 16 view features are read
// once per ray (row p / pts_per_view) instead of being materialised per sample.
//
// This file is the exact-fp32 path (every product accumulates in fp32 FFMA, like the reference's SGEMM).
// A CTA of 128 threads owns a tile of 128 points, one point per thread.  All weight matrices (37 KB, plus
// their transposes in the backward kernel) live in shared memory and are read as warp-uniform LDS.128
// broadcasts; a thread's activations live in its private column of two ping-pong shared buffers
// ([feature][thread], bank == lane, conflict-free), so layers need no barrier between them.
#include "mlp_common.cuh"

namespace hn {

__device__ __forceinline__ void load_weights_to_smem(const float* __restrict__ w, float* __restrict__ sm,
                                                     bool with_transposes) {
  for (int i = threadIdx.x; i < kWTotal; i += blockDim.x) sm[i] = packed_weight(w, i);
  if (!with_transposes) return;
  for (int i = threadIdx.x; i < kIn * kHid; i += blockDim.x) {  // Wt0[k][j] = W0[j][k]
    const int k = i / kHid, j = i % kHid;
    sm[kT0 + i] = packed_weight(w, kW0 + j * kIn + k);
  }
  for (int i = threadIdx.x; i < kHid * kH2; i += blockDim.x) {  // Wt1[k][j] = W1[j][k]
    const int k = i / kH2, j = i % kH2;
    sm[kT1 + i] = packed_weight(w, kW1 + j * kHid + k);
  }
  for (int i = threadIdx.x; i < kCinPad * kHid; i += blockDim.x) {  // Wt2[k][j] = W2p[j][k]
    const int k = i / kHid, j = i % kHid;
    sm[kT2 + i] = packed_weight(w, kW2 + j * kCinPad + k);
  }
  for (int i = threadIdx.x; i < kHid * kHid; i += blockDim.x) {  // Wt3[k][j] = W3[j][k]
    const int k = i / kHid, j = i % kHid;
    sm[kT3 + i] = packed_weight(w, kW3 + j * kHid + k);
  }
}

// One dense layer for this thread's point: out[j] = act(sum_k W[j][k] * in[k]), j < J.
// `xin` / `xout` are the thread's columns (element k at [k * kNT]); `gout`, if given, receives a copy in
// the same [feature][128] tile layout (coalesced across the warp).  `gate`: bit j clear => output forced
// to 0 (ReLU derivative in the backward pass).  Returns the mask of strictly positive pre-activations.
template <int J, int K, bool RELU, bool GATED>
__device__ __forceinline__ uint64_t dense(const float* __restrict__ W, const float* __restrict__ xin,
                                          float* __restrict__ xout, float* __restrict__ gout, uint64_t gate) {
  float in[K];
#pragma unroll
  for (int k = 0; k < K; ++k) in[k] = xin[k * kNT];
  uint64_t mask = 0;
#pragma unroll 2
  for (int j = 0; j < J; ++j) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < K; k += 8) {
      const float4 w = *reinterpret_cast<const float4*>(W + j * K + k);
      const float4 v = *reinterpret_cast<const float4*>(W + j * K + k + 4);
      a0 = fmaf(w.x, in[k], a0);
      a1 = fmaf(v.x, in[k + 4], a1);
      a0 = fmaf(w.y, in[k + 1], a0);
      a1 = fmaf(v.y, in[k + 5], a1);
      a0 = fmaf(w.z, in[k + 2], a0);
      a1 = fmaf(v.z, in[k + 6], a1);
      a0 = fmaf(w.w, in[k + 3], a0);
      a1 = fmaf(v.w, in[k + 7], a1);
    }
    float acc = a0 + a1;
    if (RELU) {
      if (acc > 0.f) mask |= (1ull << j);
      acc = fmaxf(acc, 0.f);
    }
    if (GATED) acc = ((gate >> j) & 1ull) ? acc : 0.f;
    xout[j * kNT] = acc;
    if (gout != nullptr) gout[j * kNT] = acc;
  }
  return mask;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT)
mlp_fwd_kernel(const float* __restrict__ enc, int64_t enc_stride, const float* __restrict__ views,
               int64_t views_stride, int64_t pts_per_view, const float* __restrict__ weights,
               const uint8_t* __restrict__ keep, int64_t N, float* __restrict__ out, uint32_t* __restrict__ gates) {
  extern __shared__ __align__(16) float smem[];
  float* W = smem;
  float* P = smem + kWTotal + threadIdx.x;  // ping  [64][128]
  float* Q = P + kHid * kNT;                // pong  [64][128]
  load_weights_to_smem(weights, W, false);
  __syncthreads();
  const int64_t n_tiles = (N + kNT - 1) / kNT;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * kNT + threadIdx.x;
    if (p >= N) continue;
    const float* erow = enc + p * enc_stride;
#pragma unroll
    for (int k = 0; k < kIn; ++k) P[k * kNT] = __ldg(erow + k);
    const uint64_t m1 = dense<kHid, kIn, true, false>(W + kW0, P, Q, nullptr, 0);    // h1 -> Q
    dense<kH2, kHid, false, false>(W + kW1, Q, P, nullptr, 0);   // h2 -> P[0..15]
    const float sigma = P[0];
    const float* vrow = views + (p / pts_per_view) * views_stride;
#pragma unroll
    for (int k = 0; k < kViews; ++k) Q[k * kNT] = __ldg(vrow + k);  // c = [views | geo | 0] -> Q
#pragma unroll
    for (int k = 0; k < kGeo; ++k) Q[(kViews + k) * kNT] = P[(1 + k) * kNT];
    Q[(kCinPad - 1) * kNT] = 0.f;
    const uint64_t m3 = dense<kHid, kCinPad, true, false>(W + kW2, Q, P, nullptr, 0);  // h3 -> P
    const uint64_t m4 = dense<kHid, kHid, true, false>(W + kW3, P, Q, nullptr, 0);     // h4 -> Q
    if (gates != nullptr) {
      uint2* g = reinterpret_cast<uint2*>(gates + p * 6);
      g[0] = make_uint2((uint32_t)m1, (uint32_t)(m1 >> 32));
      g[1] = make_uint2((uint32_t)m3, (uint32_t)(m3 >> 32));
      g[2] = make_uint2((uint32_t)m4, (uint32_t)(m4 >> 32));
    }
    float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < kHid; ++k) {
      const float h = Q[k * kNT];
      rgb[0] = fmaf(W[kW4 + k], h, rgb[0]);
      rgb[1] = fmaf(W[kW4 + kHid + k], h, rgb[1]);
      rgb[2] = fmaf(W[kW4 + 2 * kHid + k], h, rgb[2]);
    }
    const float s = (keep != nullptr && keep[p] == 0) ? 0.f : sigma;  // run_nerf_helpers.py:225
    reinterpret_cast<float4*>(out)[p] = make_float4(rgb[0], rgb[1], rgb[2], s);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, kernel 1: per-point deltas (the dX chain).  Activations and deltas needed by the weight
// gradients go to the caller-provided workspace, one [432][128] block per 128-point tile:
// ------------------------------------------------------------------------------------------------
constexpr int kOffH1 = 0, kOffC = 64, kOffH3 = 96, kOffH4 = 160, kOffDz1 = 224, kOffDh2 = 288, kOffDz3 = 304,
              kOffDz4 = 368, kWsRows = 432;

__global__ void __launch_bounds__(kNT)
mlp_bwd_delta_kernel(const float* __restrict__ enc, int64_t enc_stride, const float* __restrict__ views,
                     int64_t views_stride, int64_t pts_per_view, const float* __restrict__ weights,
                     const uint8_t* __restrict__ keep, const float* __restrict__ dout, int64_t N,
                     float* __restrict__ d_enc, float* __restrict__ ws) {
  extern __shared__ __align__(16) float smem[];
  float* W = smem;
  float* P = smem + kWBoth + threadIdx.x;
  float* Q = P + kHid * kNT;
  load_weights_to_smem(weights, W, true);
  __syncthreads();
  const int64_t n_tiles = (N + kNT - 1) / kNT;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * kNT + threadIdx.x;
    if (p >= N) continue;
    float* g = ws + tile * (int64_t)(kWsRows * kNT) + threadIdx.x;  // this point's workspace column
    const float* erow = enc + p * enc_stride;
#pragma unroll
    for (int k = 0; k < kIn; ++k) P[k * kNT] = __ldg(erow + k);
    const uint64_t m1 = dense<kHid, kIn, true, false>(W + kW0, P, Q, g + kOffH1 * kNT, 0);  // h1 -> Q
    dense<kH2, kHid, false, false>(W + kW1, Q, P, nullptr, 0);                              // h2 -> P
    const float* vrow = views + (p / pts_per_view) * views_stride;
#pragma unroll
    for (int k = 0; k < kViews; ++k) {
      const float v = __ldg(vrow + k);
      Q[k * kNT] = v;
      g[(kOffC + k) * kNT] = v;
    }
#pragma unroll
    for (int k = 0; k < kGeo; ++k) {
      const float v = P[(1 + k) * kNT];
      Q[(kViews + k) * kNT] = v;
      g[(kOffC + kViews + k) * kNT] = v;
    }
    Q[(kCinPad - 1) * kNT] = 0.f;
    g[(kOffC + kCinPad - 1) * kNT] = 0.f;
    const uint64_t m3 = dense<kHid, kCinPad, true, false>(W + kW2, Q, P, g + kOffH3 * kNT, 0);  // h3 -> P
    const uint64_t m4 = dense<kHid, kHid, true, false>(W + kW3, P, Q, g + kOffH4 * kNT, 0);     // h4 -> Q

    const float4 go = __ldg(reinterpret_cast<const float4*>(dout) + p);
    const float dsigma = (keep != nullptr && keep[p] == 0) ? 0.f : go.w;
    // dz4 = (W4^T drgb) . [h4 > 0]  -> P
#pragma unroll 8
    for (int k = 0; k < kHid; ++k) {
      float v = W[kW4 + k] * go.x;
      v = fmaf(W[kW4 + kHid + k], go.y, v);
      v = fmaf(W[kW4 + 2 * kHid + k], go.z, v);
      v = ((m4 >> k) & 1ull) ? v : 0.f;
      P[k * kNT] = v;
      g[(kOffDz4 + k) * kNT] = v;
    }
    dense<kHid, kHid, false, true>(W + kT3, P, Q, g + kOffDz3 * kNT, m3);  // dz3 = (W3^T dz4).[h3>0] -> Q
    // dgeo = rows 16..30 of W2^T dz3 -> P[1..15]; P[0] = dsigma  => dh2
    dense<kGeo, kHid, false, false>(W + kT2 + kViews * kHid, Q, P + kNT, g + (kOffDh2 + 1) * kNT, 0);
    P[0] = dsigma;
    g[kOffDh2 * kNT] = dsigma;
    dense<kHid, kH2, false, true>(W + kT1, P, Q, g + kOffDz1 * kNT, m1);  // dz1 = (W1^T dh2).[h1>0] -> Q
    dense<kIn, kHid, false, false>(W + kT0, Q, P, nullptr, 0);            // d_enc = W0^T dz1 -> P
    float* drow = d_enc + p * kIn;
#pragma unroll
    for (int k = 0; k < kIn; k += 4)
      *reinterpret_cast<float4*>(drow + k) =
          make_float4(P[k * kNT], P[(k + 1) * kNT], P[(k + 2) * kNT], P[(k + 3) * kNT]);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, kernel 2: weight gradients dW_l[J][K] = sum_p D_l[p][j] * A_l[p][k]
// Persistent CTAs stream 64-point half tiles of (D, A) through shared memory; every thread owns a
// TJ x TK register tile of each of the five matrices for the whole kernel and flushes once with atomics.
// ------------------------------------------------------------------------------------------------
constexpr int kChunk = 64;
constexpr int kPad = kChunk + 1;

// stage ROWS features x 64 points from the workspace tile layout ([feature][128], points contiguous)
template <int ROWS>
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int np, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < ROWS * kChunk; i += 256) {
    const int f = i / kChunk, r = i % kChunk;
    dst[f * kPad + r] = (r < np) ? __ldg(src + f * kNT + r) : 0.f;
  }
}
// stage from a row-major [N][stride] tensor (features contiguous)
template <int ROWS>
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, int64_t stride, int np,
                                           float* __restrict__ dst) {
  for (int i = threadIdx.x; i < ROWS * kChunk; i += 256) {
    const int r = i / ROWS, f = i % ROWS;
    dst[f * kPad + r] = (r < np) ? __ldg(src + r * stride + f) : 0.f;
  }
}

template <int J, int K, int TJ, int TK>
__device__ __forceinline__ void accum_pair(const float* __restrict__ sD, const float* __restrict__ sA,
                                           float (&acc)[TJ * TK]) {
  static_assert((J / TJ) * (K / TK) == 256, "tile must cover the matrix with 256 threads");
  const int tj = (threadIdx.x / (K / TK)) * TJ;
  const int tk = (threadIdx.x % (K / TK)) * TK;
#pragma unroll 4
  for (int r = 0; r < kChunk; ++r) {
    float d[TJ], a[TK];
#pragma unroll
    for (int j = 0; j < TJ; ++j) d[j] = sD[(tj + j) * kPad + r];
#pragma unroll
    for (int k = 0; k < TK; ++k) a[k] = sA[(tk + k) * kPad + r];
#pragma unroll
    for (int j = 0; j < TJ; ++j)
#pragma unroll
      for (int k = 0; k < TK; ++k) acc[j * TK + k] = fmaf(d[j], a[k], acc[j * TK + k]);
  }
}

template <int J, int K, int TJ, int TK>
__device__ __forceinline__ void flush_pair(float* __restrict__ dW, int k_valid, const float (&acc)[TJ * TK],
                                           int j_valid) {
  const int tj = (threadIdx.x / (K / TK)) * TJ;
  const int tk = (threadIdx.x % (K / TK)) * TK;
#pragma unroll
  for (int j = 0; j < TJ; ++j)
#pragma unroll
    for (int k = 0; k < TK; ++k)
      if (tj + j < j_valid && tk + k < k_valid) atomicAdd(dW + (tj + j) * k_valid + tk + k, acc[j * TK + k]);
}

__global__ void __launch_bounds__(256)
mlp_bwd_weight_kernel(const float* __restrict__ enc, int64_t enc_stride, const float* __restrict__ dout, int64_t N,
                      const float* __restrict__ ws, float* __restrict__ dweights) {
  __shared__ float sD[64 * kPad];
  __shared__ float sA[64 * kPad];
  float a0[8] = {0}, a1[4] = {0}, a2[8] = {0}, a3[16] = {0}, a4[1] = {0};
  const int64_t n_chunks = (N + kChunk - 1) / kChunk;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t p0 = ch * kChunk;
    const int np = (int)((N - p0 < kChunk) ? (N - p0) : kChunk);
    const float* t = ws + (p0 / kNT) * (int64_t)(kWsRows * kNT) + (p0 % kNT);  // half-tile base
    // dW0 = dz1^T . enc
    stage_tile<64>(t + kOffDz1 * kNT, np, sD);
    stage_rows<32>(enc + p0 * enc_stride, enc_stride, np, sA);
    __syncthreads();
    accum_pair<64, 32, 4, 2>(sD, sA, a0);
    __syncthreads();
    // dW1 = dh2^T . h1
    stage_tile<16>(t + kOffDh2 * kNT, np, sD);
    stage_tile<64>(t + kOffH1 * kNT, np, sA);
    __syncthreads();
    accum_pair<16, 64, 1, 4>(sD, sA, a1);
    __syncthreads();
    // dW2 = dz3^T . c
    stage_tile<64>(t + kOffDz3 * kNT, np, sD);
    stage_tile<32>(t + kOffC * kNT, np, sA);
    __syncthreads();
    accum_pair<64, 32, 4, 2>(sD, sA, a2);
    __syncthreads();
    // dW3 = dz4^T . h3
    stage_tile<64>(t + kOffDz4 * kNT, np, sD);
    stage_tile<64>(t + kOffH3 * kNT, np, sA);
    __syncthreads();
    accum_pair<64, 64, 4, 4>(sD, sA, a3);
    __syncthreads();
    // dW4 = drgb^T . h4   (row 3 of dout is dsigma: computed, never flushed)
    stage_rows<4>(dout + p0 * 4, 4, np, sD);
    stage_tile<64>(t + kOffH4 * kNT, np, sA);
    __syncthreads();
    accum_pair<4, 64, 1, 1>(sD, sA, a4);
    __syncthreads();
  }
  flush_pair<64, 32, 4, 2>(dweights + kG0, 32, a0, 64);
  flush_pair<16, 64, 1, 4>(dweights + kG1, 64, a1, 16);
  flush_pair<64, 32, 4, 2>(dweights + kG2, 31, a2, 64);
  flush_pair<64, 64, 4, 4>(dweights + kG3, 64, a3, 64);
  flush_pair<4, 64, 1, 1>(dweights + kG4, 64, a4, 3);
}

constexpr size_t kFwdSmem = (size_t)(kWTotal + 2 * kHid * kNT) * sizeof(float);  // 103,168 B
constexpr size_t kBwdSmem = (size_t)(kWBoth + 2 * kHid * kNT) * sizeof(float);   // 140,032 B

// tcgen05 implementation (mlp_tc.cu)
int mlp_tc_fwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
               const float* weights, const uint8_t* keep, int64_t N, float* out, uint32_t* gates, int aligned,
               cudaStream_t stream);
int64_t mlp_tc_bwd_workspace_floats(int64_t N);
int mlp_tc_bwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
               const float* weights, const uint8_t* keep, const float* dout, int64_t N, float* d_enc, float* dweights,
               float* workspace, int aligned, cudaStream_t stream);
int64_t mlp_tc_bwd_fused_workspace_bytes();
int mlp_tc_bwd_fused(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
                     const float* weights, const uint8_t* keep, const uint32_t* gates, const float* dout, int64_t N,
                     float* d_enc, float* dweights, float* workspace, int aligned, cudaStream_t stream);
int g_mlp_impl = 1;      // 1 = tcgen05 tensor-core MLP (mlp_tc.cu / mlp_tc_bwd.cu, default), 0 = FFMA fp32 (this file)
int g_mlp_bwd_impl = 1;  // with mlp_impl = 1: 1 = fused bf16x2 backward (mlp_tc_bwd.cu, default), 0 = two-kernel 3xTF32

static inline bool rows16(const float* p, int64_t stride) {
  return ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) && (stride % 4 == 0);
}

static int ensure_smem_optin() {
  static thread_local int done_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, "cudaGetDevice");
  if (done_dev == dev) return 0;
  e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_fwd_kernel)");
  e = cudaFuncSetAttribute(mlp_bwd_delta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_bwd_delta_kernel)");
  done_dev = dev;
  return 0;
}

}  // namespace hn

extern "C" {

int64_t hn_mlp_bwd_workspace_bytes(int64_t N) {
  if (N <= 0) return 0;
  // sized for the implementation selected now (hn_set_tuning); query again after changing it
  if (hn::g_mlp_impl == 1 && hn::g_mlp_bwd_impl == 1) return hn::mlp_tc_bwd_fused_workspace_bytes();
  const int64_t tiles = (N + hn::kNT - 1) / hn::kNT;
  const int64_t ffma = tiles * hn::kWsRows * hn::kNT;
  const int64_t tcw = hn::mlp_tc_bwd_workspace_floats(N);
  return (ffma > tcw ? ffma : tcw) * (int64_t)sizeof(float);
}

int hn_mlp_fwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
               const float* weights, const uint8_t* keep, int64_t N, float* out, uint32_t* gates, void* stream) {
  HN_REQUIRE(N >= 0, "hn_mlp_fwd: negative N");
  HN_REQUIRE(pts_per_view >= 1, "hn_mlp_fwd: pts_per_view must be >= 1");
  HN_REQUIRE(enc_stride >= 32 && views_stride >= 0, "hn_mlp_fwd: bad row stride");
  if (N == 0) return 0;
  HN_REQUIRE(enc && views && weights && out, "hn_mlp_fwd: null pointer");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "hn_mlp_fwd: out must be 16-byte aligned");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(gates) & 7u) == 0, "hn_mlp_fwd: gates must be 8-byte aligned");
  if (hn::g_mlp_impl == 1)
    return hn::mlp_tc_fwd(enc, enc_stride, views, views_stride, pts_per_view, weights, keep, N, out, gates,
                          hn::rows16(enc, enc_stride) ? 1 : 0, (cudaStream_t)stream);
  int rc = hn::ensure_smem_optin();
  if (rc) return rc;
  const int64_t tiles = (N + hn::kNT - 1) / hn::kNT;
  const int64_t cap = (int64_t)hn::sm_count() * 2;  // 2 CTAs of 101 KB fit per SM
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  hn::mlp_fwd_kernel<<<grid, hn::kNT, hn::kFwdSmem, (cudaStream_t)stream>>>(enc, enc_stride, views, views_stride,
                                                                           pts_per_view, weights, keep, N, out, gates);
  return hn::check_launch("mlp_fwd_kernel");
}

int hn_mlp_bwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
               const float* weights, const uint8_t* keep, const uint32_t* gates, const float* dout, int64_t N,
               float* d_enc, float* dweights, float* workspace, void* stream) {
  HN_REQUIRE(N >= 0, "hn_mlp_bwd: negative N");
  HN_REQUIRE(pts_per_view >= 1, "hn_mlp_bwd: pts_per_view must be >= 1");
  HN_REQUIRE(enc_stride >= 32 && views_stride >= 0, "hn_mlp_bwd: bad row stride");
  if (N == 0) return 0;
  HN_REQUIRE(enc && views && weights && dout && d_enc && dweights && workspace, "hn_mlp_bwd: null pointer");
  HN_REQUIRE(((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(d_enc)) & 15u) == 0,
             "hn_mlp_bwd: dout and d_enc must be 16-byte aligned");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(gates) & 7u) == 0, "hn_mlp_bwd: gates must be 8-byte aligned");
  if (hn::g_mlp_impl == 1 && hn::g_mlp_bwd_impl == 1)
    return hn::mlp_tc_bwd_fused(enc, enc_stride, views, views_stride, pts_per_view, weights, keep, gates, dout, N, d_enc,
                                dweights, workspace, hn::rows16(enc, enc_stride) ? 1 : 0, (cudaStream_t)stream);
  if (hn::g_mlp_impl == 1)
    return hn::mlp_tc_bwd(enc, enc_stride, views, views_stride, pts_per_view, weights, keep, dout, N, d_enc, dweights,
                          workspace, hn::rows16(enc, enc_stride) ? 1 : 0, (cudaStream_t)stream);
  int rc = hn::ensure_smem_optin();
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  {
    const int64_t tiles = (N + hn::kNT - 1) / hn::kNT;
    const int64_t cap = (int64_t)hn::sm_count();  // 137 KB per CTA: one per SM
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    hn::mlp_bwd_delta_kernel<<<grid, hn::kNT, hn::kBwdSmem, s>>>(enc, enc_stride, views, views_stride, pts_per_view,
                                                                 weights, keep, dout, N, d_enc, workspace);
    rc = hn::check_launch("mlp_bwd_delta_kernel");
    if (rc) return rc;
  }
  {
    const int64_t n_chunks = (N + hn::kChunk - 1) / hn::kChunk;
    const int64_t cap = (int64_t)hn::sm_count() * 4;
    const unsigned grid = (unsigned)(n_chunks < cap ? n_chunks : cap);
    hn::mlp_bwd_weight_kernel<<<grid, 256, 0, s>>>(enc, enc_stride, dout, N, workspace, dweights);
    return hn::check_launch("mlp_bwd_weight_kernel");
  }
}

}  // extern "C"
