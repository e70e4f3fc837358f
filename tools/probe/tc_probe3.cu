// tc_probe3.cu -- validates, on a B200, the tcgen05 encodings the fused MLP backward kernel relies on:
//   T1  kind::f16 (bf16) TS-form MMA: A packed two K elements per TMEM column (element 2c in the low half),
//       B = canonical no-swizzle K-major bf16 image; M = 128, N in {64, 16, 32}, K in {16, 32, 64}
//   T2  kind::f16 SS-form MMA with BOTH operands MN-major (the contraction index = points is the slow index):
//       A = [point][64 features] 128-byte lines, 128-byte swizzle; B = same line layout with N in {64,32,16,8}
//       valid features, or N = 8 in the dense 16-byte-per-point no-swizzle layout; M = 64; accumulator at TMEM
//       lane offset 0 and 16 (the interleaved placement of two M = 64 accumulators in the same columns)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hashnerf-pytorch_b200/csrc -o tools/probe/tc_probe3 tools/probe/tc_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "tc05.cuh"
using namespace hn::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

static float bf16r(float a) {  // round to nearest even onto the bf16 grid
  uint32_t u; memcpy(&u, &a, 4);
  u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u;
  memcpy(&a, &u, 4); return a;
}
__device__ __forceinline__ uint16_t bf16_bits(float a) {
  uint32_t u = __float_as_uint(a);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// ---- T1: D[128 x N] = A[128 x K] . W[N x K]^T ; order: 0 = element 2c in the low half, 1 = in the high half
__global__ void __launch_bounds__(128) probe_ts(const float* A, const float* W, float* D, int K, int N, int order) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint16_t* Ws = reinterpret_cast<uint16_t*>(smem);
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < N * K; i += 128) Ws[canon16(i / K, i % K, K)] = bf16_bits(W[i]);
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
  // A operand at columns 64.. : K/2 words
  for (int c0 = 0; c0 < K / 2; c0 += 8) {
    uint32_t w[8];
    for (int i = 0; i < 8; ++i) {
      const uint32_t e0 = bf16_bits(A[t * K + 2 * (c0 + i)]), e1 = bf16_bits(A[t * K + 2 * (c0 + i) + 1]);
      w[i] = order ? ((e0 << 16) | e1) : ((e1 << 16) | e0);
    }
    tmem_st8(row + 64 + c0, w);
  }
  wait_st();
  fence_before_sync();
  __syncthreads();
  if (t == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t lbo = 128, sbo = (K / 8) * 128;
    for (int s = 0; s < K / 16; ++s)
      umma_ts_bf16(tmem, tmem + 64 + 8 * s, make_sdesc(smem_u32(Ws) + s * 256, lbo, sbo), idesc, s ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c = 0; c < N; c += 8) {
    float v[8];
    tmem_ld8(row + c, v);
    for (int i = 0; i < 8; ++i) D[t * N + c + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- T2: D[64 x N] = sum_p X[p][m] * Y[p][n], P = 128 points.  ylayout: 0 = 128-byte swizzled lines, 1 = dense
// 16-byte rows (N = 8 only).  lane_off: 0 or 16.
__global__ void __launch_bounds__(128) probe_mn(const float* X, const float* Y, float* D, int N, int ylayout, int lane_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint16_t* Xs = reinterpret_cast<uint16_t*>(smem);              // 128 points x 128 B
  uint16_t* Ys = reinterpret_cast<uint16_t*>(smem + 128 * 128);  // 128 points x 128 B (or x 16 B)
  const int t = threadIdx.x, warp = t >> 5;
  // thread t = point t writes its own line: chunk j (8 features) at chunk position j ^ (t & 7)
  for (int f = 0; f < 64; ++f) Xs[t * 64 + (((f >> 3) ^ (t & 7)) << 3) + (f & 7)] = bf16_bits(X[t * 64 + f]);
  if (ylayout == 0) {
    for (int f = 0; f < 64; ++f) Ys[t * 64 + (((f >> 3) ^ (t & 7)) << 3) + (f & 7)] = f < N ? bf16_bits(Y[t * 64 + f]) : (uint16_t)0x7FC0;  // NaN outside
  } else {
    for (int f = 0; f < 8; ++f) Ys[t * 8 + f] = bf16_bits(Y[t * 64 + f]);
  }
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  {  // poison the accumulator columns so that untouched lanes are recognisable
    uint32_t w[8];
    for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(-777.f);
    for (int c = 0; c < 64; c += 8) tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + c, w);
    wait_st();
  }
  fence_before_sync();
  __syncthreads();
  if (t == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(64, N, 1, 1);
    for (int s = 0; s < 128 / 16; ++s) {
      const uint64_t ad = make_sdesc_mn_sw128(smem_u32(Xs) + s * 2048);
      const uint64_t bd = ylayout == 0 ? make_sdesc_mn_sw128(smem_u32(Ys) + s * 2048) : make_sdesc_mn_n8(smem_u32(Ys) + s * 256);
      umma_ss_bf16(tmem + ((uint32_t)lane_off << 16), ad, bd, idesc, s ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c = 0; c < 64; c += 8) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 8; ++i) D[t * 64 + c + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  int fails = 0;
  // ---------------- T1
  struct C1 { int K, N; } c1[] = {{32, 64}, {64, 64}, {64, 16}, {16, 64}, {64, 32}};
  for (auto c : c1) for (int order = 0; order < 2; ++order) {
    const int K = c.K, N = c.N;
    std::vector<float> A(128 * K), W(N * K), D(128 * N);
    srand(K * 131 + N * 7);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *dA, *dW, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)N * K * 2 + 1024;
    probe_ts<<<1, 128, smem>>>(dA, dW, dD, K, N, order);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)bf16r(A[m * K + k]) * (double)bf16r(W[n * K + k]);
      max_err = fmax(max_err, fabs(ref - D[m * N + n]));
    }
    const bool ok = max_err < 1e-4;
    printf("T1 TS bf16 K=%2d N=%2d order=%d  max_abs_err=%.3e %s\n", K, N, order, max_err, ok ? "OK" : "bad");
    if (order == 0) fails += !ok;
    cudaFree(dA); cudaFree(dW); cudaFree(dD);
  }
  // ---------------- T2
  struct C2 { int N, ylayout, lane_off; } c2[] = {{64, 0, 0}, {32, 0, 0}, {16, 0, 0}, {8, 0, 0}, {8, 1, 0}, {64, 0, 16}, {32, 0, 16}, {8, 1, 16}};
  CK(cudaFuncSetAttribute(probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  for (auto c : c2) {
    std::vector<float> X(128 * 64), Y(128 * 64), D(128 * 64);
    srand(c.N * 17 + c.ylayout * 3 + c.lane_off);
    for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : Y) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *dX, *dY, *dD;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dY, Y.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice));
    probe_mn<<<1, 128, 34 * 1024>>>(dX, dY, dD, c.N, c.ylayout, c.lane_off);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0; int untouched_ok = 1;
    for (int m = 0; m < 64; ++m) {
      const int lane = (m % 16) + 32 * (m / 16) + c.lane_off;
      for (int n = 0; n < c.N; ++n) {
        double ref = 0;
        for (int p = 0; p < 128; ++p) ref += (double)bf16r(X[p * 64 + m]) * (double)bf16r(Y[p * 64 + n]);
        const double e = fabs(ref - D[lane * 64 + n]);
        max_err = fmax(max_err, std::isnan(e) ? 1e30 : e);
      }
      // the other half of each 32-lane quadrant, and columns >= N, must be untouched
      const int other = (m % 16) + 32 * (m / 16) + (16 - c.lane_off);
      for (int n = 0; n < 64; ++n) if (D[other * 64 + n] != -777.f) untouched_ok = 0;
      for (int n = c.N; n < 64; ++n) if (D[lane * 64 + n] != -777.f) untouched_ok = 0;
    }
    const bool ok = max_err < 2e-4 && untouched_ok;
    printf("T2 SS MN-major N=%2d ylayout=%d lane_off=%2d  max_abs_err=%.3e untouched=%d %s\n", c.N, c.ylayout, c.lane_off, max_err, untouched_ok, ok ? "OK" : "bad");
    fails += !ok;
    cudaFree(dX); cudaFree(dY); cudaFree(dD);
  }
  printf(fails ? "PROBE3 FAILED (%d)\n" : "PROBE3 PASSED\n", fails);
  return fails ? 1 : 0;
}
