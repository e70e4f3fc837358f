// rays.cu -- ray generation and ray-batch packing ("next" row 3).
//
//   hn_get_rays   pinhole camera rays of a whole image        (reference ray_util.py:62-80)
//   hn_pack_rays  [o | d | near | far | d_view/|d_view|]      (reference run_nerf_helpers.py:344-366)
//
// Both replace chains of ~8-10 small ATen launches per training iteration; arithmetic is in the reference's
// fp32 rounding order.
#include "common.cuh"

namespace hn {

__global__ void __launch_bounds__(256)
get_rays_kernel(int H, int W, float fx, float fy, float cx, float cy, const float* __restrict__ c2w, int64_t row_stride,
                float* __restrict__ rays_d) {
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (int64_t)H * W) return;
  const int col = (int)(pix % W), row = (int)(pix / W);
  // dirs = ((i - cx)/fx, -(j - cy)/fy, -1)                                               ray_util.py:75
  const float d0 = __fdiv_rn(__fsub_rn((float)col, cx), fx);
  const float d1 = __fdiv_rn(-__fsub_rn((float)row, cy), fy);
  const float d2 = -1.f;
  // rays_d[c] = sum_k dirs[k] * c2w[c][k]                                                 :78
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* m = c2w + c * row_stride;
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(d0, __ldg(m)), __fmul_rn(d1, __ldg(m + 1))), __fmul_rn(d2, __ldg(m + 2)));
    rays_d[pix * 3 + c] = s;
  }
}

__global__ void __launch_bounds__(256)
pack_rays_kernel(const float* __restrict__ o, int64_t o_stride, const float* __restrict__ d, int64_t d_stride,
                 const float* __restrict__ vd, int64_t vd_stride, float near, float far, int64_t R, int width,
                 float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float* dst = out + r * width;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dst[c] = __ldg(o + r * o_stride + c);
    dst[3 + c] = __ldg(d + r * d_stride + c);
  }
  dst[6] = near;
  dst[7] = far;
  if (vd != nullptr) {  // viewdirs / ||viewdirs||                                         run_nerf_helpers.py:350
    const float x = __ldg(vd + r * vd_stride), y = __ldg(vd + r * vd_stride + 1), z = __ldg(vd + r * vd_stride + 2);
    const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    dst[8] = __fdiv_rn(x, n);
    dst[9] = __fdiv_rn(y, n);
    dst[10] = __fdiv_rn(z, n);
  }
}

}  // namespace hn

extern "C" {

int hn_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int64_t c2w_row_stride,
                float* rays_d, void* stream) {
  HN_REQUIRE(H >= 1 && W >= 1, "hn_get_rays: bad image size");
  HN_REQUIRE(c2w && rays_d, "hn_get_rays: null pointer");
  const int64_t n = (int64_t)H * W;
  hn::get_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(H, W, fx, fy, cx, cy, c2w,
                                                                                     c2w_row_stride, rays_d);
  return hn::check_launch("get_rays_kernel");
}

int hn_pack_rays(const float* rays_o, int64_t o_stride, const float* rays_d, int64_t d_stride, const float* viewdirs,
                 int64_t vd_stride, float near, float far, int64_t R, float* out, void* stream) {
  HN_REQUIRE(R >= 0, "hn_pack_rays: negative R");
  if (R == 0) return 0;
  HN_REQUIRE(rays_o && rays_d && out, "hn_pack_rays: null pointer");
  hn::pack_rays_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, o_stride, rays_d, d_stride, viewdirs, vd_stride, near, far, R, viewdirs ? 11 : 8, out);
  return hn::check_launch("pack_rays_kernel");
}

}  // extern "C"
