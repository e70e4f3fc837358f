// tc_probe.cu -- standalone validation of the tcgen05 / TMEM mechanics used by the fused MLP:
// smem descriptors (no-swizzle canonical layouts, K-major and MN-major), instruction descriptor for
// kind::tf32, TMEM alloc / ld, mbarrier commit, and the 3xTF32 split.   nvcc -arch=sm_100a tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity));
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// canonical no-swizzle K-major placement of element (r, k) of an [R x K] fp32 operand
__host__ __device__ inline int canon(int r, int k, int K) { return ((r >> 3) * (K >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ void split(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  lo = a - hi;
}

// mode 0: D[128 x N] = A[128 x K] . B[N x K]^T with single TF32;  mode 1: same with 3xTF32
// mode 2: C[64 x 64] = sum_p X[p][j] * Y[p][k], X = A (128 x 64, K-major stored, read MN-major), Y = B (128 x 64), 3xTF32
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int K, int N, int mode) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int RA = 128, RB = (mode == 2) ? 128 : N;
  float* Ahi = smem;
  float* Alo = Ahi + RA * K;
  float* Bhi = Alo + RA * K;
  float* Blo = Bhi + RB * K;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < RA * K; i += 128) {
    const int r = i / K, k = i % K;
    float hi, lo;
    split(A[i], hi, lo);
    Ahi[canon(r, k, K)] = hi;
    Alo[canon(r, k, K)] = lo;
  }
  for (int i = t; i < RB * K; i += 128) {
    const int r = i / K, k = i % K;
    float hi, lo;
    split(B[i], hi, lo);
    Bhi[canon(r, k, K)] = hi;
    Blo[canon(r, k, K)] = lo;
  }
  if (t == 0) mbar_init(&bar, 1);
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  fence_async_smem();
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    if (mode < 2) {
      const uint32_t idesc = make_idesc(128, N, 0, 0);
      const uint32_t lbo = 128, sbo = (K / 4) * 128;
      int first = 1;
      for (int pass = 0; pass < (mode == 1 ? 3 : 1); ++pass) {
        const float* a = (pass == 2) ? Alo : Ahi;
        const float* b = (pass == 1) ? Blo : Bhi;
        for (int s = 0; s < K / 8; ++s) {
          umma_tf32(tmem, make_sdesc(smem_u32(a) + s * 256, lbo, sbo), make_sdesc(smem_u32(b) + s * 256, lbo, sbo), idesc,
                    first ? 0u : 1u);
          first = 0;
        }
      }
    } else {
      // MN-major reads of the K-major stored buffers: MN chunk stride (SBO) = 128 B, K-group stride = (K/4)*128 B
      const uint32_t idesc = make_idesc(64, 64, 1, 1);
      const uint32_t kgroup = (K / 4) * 128;
      int first = 1;
      for (int pass = 0; pass < 3; ++pass) {
        const float* a = (pass == 2) ? Alo : Ahi;
        const float* b = (pass == 1) ? Blo : Bhi;
        for (int s = 0; s < 128 / 8; ++s) {
          umma_tf32(tmem, make_sdesc(smem_u32(a) + s * kgroup, kgroup, 128), make_sdesc(smem_u32(b) + s * kgroup, kgroup, 128),
                    idesc, first ? 0u : 1u);
          first = 0;
        }
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (mode < 2) {
    for (int c = 0; c < N; c += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      for (int i = 0; i < 8; ++i) D[t * N + c + i] = v[i];
    }
  } else {
    // dump all 128 lanes x 64 columns; the host works out where the rows went
    for (int c = 0; c < 64; c += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      for (int i = 0; i < 8; ++i) D[t * 64 + c + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float tf32_trunc(float a) { uint32_t u; memcpy(&u, &a, 4); u &= 0xFFFFE000u; memcpy(&a, &u, 4); return a; }

int main() {
  int fails = 0;
  struct Case { int K, N, mode; } cases[] = {{32, 64, 0}, {64, 64, 0}, {64, 16, 0}, {64, 8, 0}, {32, 64, 1}, {64, 64, 1}, {64, 16, 1},
                                             {64, 8, 1}, {16, 64, 1}, {8, 64, 1}, {64, 32, 1}, {64, 64, 2}};
  for (auto c : cases) {
    const int RA = 128, RB = (c.mode == 2) ? 128 : c.N, K = c.K;
    std::vector<float> A(RA * K), B(RB * K);
    srand(c.K * 131 + c.N * 7 + c.mode);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    const int outM = 128, outN = (c.mode == 2) ? 64 : c.N;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, outM * outN * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, outM * outN * 4));
    const size_t smem = (size_t)(2 * RA * K + 2 * RB * K) * 4;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 128, smem>>>(dA, dB, dD, K, c.N, c.mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(outM * outN);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    if (c.mode == 2) {
      // reference C[j][k]
      std::vector<double> C(64 * 64);
      for (int j = 0; j < 64; ++j) for (int k = 0; k < 64; ++k) { double r = 0; for (int p = 0; p < 128; ++p) r += (double)A[p * K + j] * (double)B[p * K + k]; C[j * 64 + k] = r; }
      for (int lane = 0; lane < 128; ++lane) {
        int best = -1; double best_err = 1e30; int transposed = 0;
        for (int j = 0; j < 64; ++j) {
          double e = 0, et = 0;
          for (int k = 0; k < 64; ++k) { e = fmax(e, fabs(C[j * 64 + k] - D[lane * 64 + k])); et = fmax(et, fabs(C[k * 64 + j] - D[lane * 64 + k])); }
          if (e < best_err) { best_err = e; best = j; transposed = 0; }
          if (et < best_err) { best_err = et; best = j; transposed = 1; }
        }
        if (lane < 4 || lane % 16 == 0 || best_err < 1e-3)
          printf("  lane %3d: best row %2d%s err %.3e  first vals %.4f %.4f %.4f (ref row0 %.4f %.4f %.4f)\n", lane, best, transposed ? "T" : "", best_err,
                 D[lane * 64], D[lane * 64 + 1], D[lane * 64 + 2], C[0], C[1], C[2]);
      }
      continue;
    }
    for (int m = 0; m < outM; ++m)
      for (int n = 0; n < outN; ++n) {
        double ref = 0;
        if (c.mode == 0) for (int k = 0; k < K; ++k) ref += (double)tf32_trunc(A[m * K + k]) * (double)tf32_trunc(B[n * K + k]);
        else if (c.mode == 1) for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
        else for (int p = 0; p < 128; ++p) ref += (double)A[p * K + m] * (double)B[p * K + n];
        max_err = fmax(max_err, fabs(ref - (double)D[m * outN + n]));
        max_ref = fmax(max_ref, fabs(ref));
      }
    const double tol = (c.mode == 0) ? 2e-6 : 4e-6;
    const bool ok = max_err <= tol * fmax(1.0, max_ref);
    printf("K=%2d N=%2d mode=%d  max_abs_err=%.3e  max_ref=%.3f  %s\n", c.K, c.N, c.mode, max_err, max_ref, ok ? "OK" : "FAIL");
    fails += !ok;
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  printf(fails ? "PROBE FAILED (%d)\n" : "PROBE PASSED\n", fails);
  return fails ? 1 : 0;
}
