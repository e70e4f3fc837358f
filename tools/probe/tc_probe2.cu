// tc_probe2.cu -- which (LBO, SBO) convention do MN-major no-swizzle tf32 operands take?
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity));
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ inline int canon(int r, int k, int K) { return ((r >> 3) * (K >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

// D[128 x 64] = sum_{k<64} A(m,k) * B(n,k).  a_mn: A given as At[k][m] (64 x 128) stored canon(k, m, 128); else A[m][k] canon(m,k,64)
// b_mn: B given as Bt[k][n] (64 x 64) stored canon(k, n, 64); else B[n][k] canon(n, k, 64).   conv: 0 = (LBO = k-group stride, SBO = MN-chunk stride), 1 = swapped
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int a_mn, int b_mn, int conv) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  float* As = smem;            // 128*64
  float* Bs = smem + 128 * 64; // 64*64
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 128 * 64; i += 128) {
    const int m = i / 64, k = i % 64;
    const float v = __uint_as_float(__float_as_uint(A[i]) & 0xFFFFE000u);
    As[a_mn ? canon(k, m, 128) : canon(m, k, 64)] = v;
  }
  for (int i = t; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const float v = __uint_as_float(__float_as_uint(B[i]) & 0xFFFFE000u);
    Bs[b_mn ? canon(k, n, 64) : canon(n, k, 64)] = v;
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
    const uint32_t idesc = make_idesc(128, 64, a_mn, b_mn);
    for (int s = 0; s < 8; ++s) {
      uint64_t ad, bd;
      if (!a_mn) ad = make_sdesc(smem_u32(As) + s * 256, 128, 16 * 128);
      else {
        const uint32_t kg = 32 * 128, mn = 128;  // k-group stride, MN-chunk stride of the (64 x 128) buffer
        ad = conv ? make_sdesc(smem_u32(As) + s * kg, mn, kg) : make_sdesc(smem_u32(As) + s * kg, kg, mn);
      }
      if (!b_mn) bd = make_sdesc(smem_u32(Bs) + s * 256, 128, 16 * 128);
      else {
        const uint32_t kg = 16 * 128, mn = 128;
        bd = conv ? make_sdesc(smem_u32(Bs) + s * kg, mn, kg) : make_sdesc(smem_u32(Bs) + s * kg, kg, mn);
      }
      umma_tf32(tmem, ad, bd, idesc, s ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < 64; c += 8) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 8; ++i) D[t * 64 + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}
static float tf32_trunc(float a) { uint32_t u; memcpy(&u, &a, 4); u &= 0xFFFFE000u; memcpy(&a, &u, 4); return a; }
int main() {
  std::vector<float> A(128 * 64), B(64 * 64), D(128 * 64);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (128 * 64 + 64 * 64) * 4;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int a_mn = 0; a_mn < 2; ++a_mn) for (int b_mn = 0; b_mn < 2; ++b_mn) for (int conv = 0; conv < 2; ++conv) {
    if (!a_mn && !b_mn && conv) continue;
    CK(cudaMemset(dD, 0, D.size() * 4));
    probe<<<1, 128, smem>>>(dA, dB, dD, a_mn, b_mn, conv);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)tf32_trunc(A[m * 64 + k]) * (double)tf32_trunc(B[n * 64 + k]);
      max_err = fmax(max_err, fabs(ref - (double)D[m * 64 + n])); max_ref = fmax(max_ref, fabs(ref));
    }
    printf("a_mn=%d b_mn=%d conv=%d  max_abs_err=%.3e (max_ref %.3f) %s\n", a_mn, b_mn, conv, max_err, max_ref, max_err < 1e-4 ? "OK" : "bad");
  }
  return 0;
}
