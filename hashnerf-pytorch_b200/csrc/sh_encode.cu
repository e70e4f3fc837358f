// sh_encode.cu -- real spherical-harmonics direction encoding, degree 1..5.
//
// Replaces SHEncoder.forward (reference embedding/spherical_harmonic.py:65-103; constants :13-41).
// Arithmetic follows the reference op for op (python-float constants are rounded to fp32 where they
// multiply a tensor; products evaluate left to right; no FMA contraction) so the result is bit-exact.
#include "common.cuh"

namespace hn {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }

__device__ __forceinline__ void sh_eval(float x, float y, float z, int degree, float* __restrict__ o) {
  constexpr float C0 = 0.28209479177387814f;
  constexpr float C1 = 0.4886025119029199f;
  o[0] = C0;
  if (degree <= 1) return;
  o[1] = mul(-C1, y);
  o[2] = mul(C1, z);
  o[3] = mul(-C1, x);
  if (degree <= 2) return;
  const float xx = mul(x, x), yy = mul(y, y), zz = mul(z, z);
  const float xy = mul(x, y), yz = mul(y, z), xz = mul(x, z);
  o[4] = mul(1.0925484305920792f, xy);
  o[5] = mul(-1.0925484305920792f, yz);
  o[6] = mul(0.31539156525252005f, sub(sub(mul(2.0f, zz), xx), yy));
  o[7] = mul(-1.0925484305920792f, xz);
  o[8] = mul(0.5462742152960396f, sub(xx, yy));
  if (degree <= 3) return;
  o[9] = mul(mul(-0.5900435899266435f, y), sub(mul(3.f, xx), yy));
  o[10] = mul(mul(2.890611442640554f, xy), z);
  o[11] = mul(mul(-0.4570457994644658f, y), sub(sub(mul(4.f, zz), xx), yy));
  o[12] = mul(mul(0.3731763325901154f, z), sub(sub(mul(2.f, zz), mul(3.f, xx)), mul(3.f, yy)));
  o[13] = mul(mul(-0.4570457994644658f, x), sub(sub(mul(4.f, zz), xx), yy));
  o[14] = mul(mul(1.445305721320277f, z), sub(xx, yy));
  o[15] = mul(mul(-0.5900435899266435f, x), sub(xx, mul(3.f, yy)));
  if (degree <= 4) return;
  o[16] = mul(mul(2.5033429417967046f, xy), sub(xx, yy));
  o[17] = mul(mul(-1.7701307697799304f, yz), sub(mul(3.f, xx), yy));
  o[18] = mul(mul(0.9461746957575601f, xy), sub(mul(7.f, zz), 1.f));
  o[19] = mul(mul(-0.6690465435572892f, yz), sub(mul(7.f, zz), 3.f));
  o[20] = mul(0.10578554691520431f, add(mul(zz, sub(mul(35.f, zz), 30.f)), 3.f));
  o[21] = mul(mul(-0.6690465435572892f, xz), sub(mul(7.f, zz), 3.f));
  o[22] = mul(mul(0.47308734787878004f, sub(xx, yy)), sub(mul(7.f, zz), 1.f));
  o[23] = mul(mul(-1.7701307697799304f, xz), sub(xx, mul(3.f, yy)));
  o[24] = mul(0.6258357354491761f, sub(mul(xx, sub(xx, mul(3.f, yy))), mul(yy, sub(mul(3.f, xx), yy))));
}

__global__ void __launch_bounds__(256)
sh_encode_kernel(const float* __restrict__ dirs, int64_t N, int degree, float* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  float o[25];
  sh_eval(__ldg(dirs + p * 3), __ldg(dirs + p * 3 + 1), __ldg(dirs + p * 3 + 2), degree, o);
  const int n = degree * degree;
  float* dst = out + p * n;
  if (n == 16) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      reinterpret_cast<float4*>(dst)[v] = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 25; ++i)
      if (i < n) dst[i] = o[i];
  }
}

}  // namespace hn

extern "C" int hn_sh_encode(const float* dirs, int64_t N, int degree, float* out, void* stream) {
  HN_REQUIRE(degree >= 1 && degree <= 5, "hn_sh_encode: degree must be in [1,5]");
  HN_REQUIRE(N >= 0, "hn_sh_encode: negative N");
  if (N == 0) return 0;
  HN_REQUIRE(dirs && out, "hn_sh_encode: null pointer");
  hn::sh_encode_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dirs, N, degree, out);
  return hn::check_launch("sh_encode_kernel");
}
