// rays.cu -- ray generation and ray-batch packing ("next" row 3).
//
//   hn_get_rays   pinhole camera rays of a whole image        (reference ray_util.py:62-80)
//   hn_pack_rays  [o | d | near | far | d_view/|d_view|]      (reference run_nerf_helpers.py:344-366)
//
//   hn_ndc_rays   forward-facing NDC warp of (o, d)             (reference ray_util.py:96-142)
//   hn_sample_rays  opt-in training batcher: N_rand distinct pixels of one image -> packed rays + targets
//                   (replaces run_nerf.py:576-605: whole-image get_rays, host np.random.choice, index upload,
//                   gather of rays and targets, and the packing of run_nerf_helpers.py:344-366)
//
// All replace chains of ~8-10 small ATen launches per training iteration; arithmetic is in the reference's
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

// ray_util.py:119-140, one ray per thread, every operation an individually rounded fp32 op in the reference's order.
// sx = -1/(W/(2 focal)), sy = -1/(H/(2 focal)) and two_near = 2 near are formed in double on the host, as the
// reference's python scalars are, and rounded to fp32 where ATen would round them.
__global__ void __launch_bounds__(256)
ndc_rays_kernel(const float* __restrict__ o, int64_t o_stride, const float* __restrict__ d, int64_t d_stride, int64_t R,
                float near, float two_near, float sx, float sy, float* __restrict__ out_o, float* __restrict__ out_d) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float ox = __ldg(o + r * o_stride), oy = __ldg(o + r * o_stride + 1), oz = __ldg(o + r * o_stride + 2);
  const float dx = __ldg(d + r * d_stride), dy = __ldg(d + r * d_stride + 1), dz = __ldg(d + r * d_stride + 2);
  const float t = __fdiv_rn(-__fadd_rn(near, oz), dz);                                  // :119
  ox = __fadd_rn(ox, __fmul_rn(t, dx));                                                 // :120
  oy = __fadd_rn(oy, __fmul_rn(t, dy));
  oz = __fadd_rn(oz, __fmul_rn(t, dz));
  const float ox_oz = __fdiv_rn(ox, oz), oy_oz = __fdiv_rn(oy, oz);                     // :124-125
  const float o0 = __fmul_rn(sx, ox_oz), o1 = __fmul_rn(sy, oy_oz);                     // :129-130
  const float o2 = __fadd_rn(1.f, __fdiv_rn(two_near, oz));                             // :131
  const float d0 = __fmul_rn(sx, __fsub_rn(__fdiv_rn(dx, dz), ox_oz));                  // :134
  const float d1 = __fmul_rn(sy, __fsub_rn(__fdiv_rn(dy, dz), oy_oz));
  const float d2 = __fsub_rn(1.f, o2);                                                  // :136
  out_o[r * 3] = o0;
  out_o[r * 3 + 1] = o1;
  out_o[r * 3 + 2] = o2;
  out_d[r * 3] = d0;
  out_d[r * 3 + 1] = d1;
  out_d[r * 3 + 2] = d2;
}

// ---- on-device training batcher ------------------------------------------------------------------------------
// N_rand DISTINCT pixels of a win_h x win_w window (the whole image, or the centre crop of run_nerf.py:586-596):
// pixel k of the batch is the image of k under a keyed pseudo-random permutation of [0, win_h * win_w) -- a 4-round
// Feistel network over the next power of four, cycle-walked back into range -- so the first N_rand values are a
// sample WITHOUT replacement, like np.random.choice(..., replace=False) at :600.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint32_t feistel_perm(uint32_t k, uint32_t n, uint32_t half_bits, uint32_t seed) {
  const uint32_t half_mask = (1u << half_bits) - 1u;
  uint32_t v = k;
  do {
    uint32_t l = v >> half_bits, r = v & half_mask;
#pragma unroll
    for (uint32_t round = 0; round < 4; ++round) {
      const uint32_t f = mix32(r ^ (seed + 0x9e3779b9u * (round + 1u))) & half_mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    v = (l << half_bits) | r;
  } while (v >= n);  // cycle walking: the permutation of [0, 4^half_bits) restricted to [0, n) is a permutation
  return v;
}

// step = {image index, seed, row0, col0, win_h, win_w} in DEVICE memory: the launch can live inside a CUDA graph
__global__ void __launch_bounds__(256)
sample_rays_kernel(const float* __restrict__ images, int C, const float* __restrict__ poses, int H, int W, float fx,
                   float fy, float cx, float cy, float near, float far, const int32_t* __restrict__ step, int64_t n_rand,
                   float* __restrict__ rays, float* __restrict__ target, int32_t* __restrict__ pix) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_rand) return;
  const int img = __ldg(step), row0 = __ldg(step + 2), col0 = __ldg(step + 3), win_h = __ldg(step + 4),
            win_w = __ldg(step + 5);
  const uint32_t seed = (uint32_t)__ldg(step + 1);
  const uint32_t n = (uint32_t)win_h * (uint32_t)win_w;
  uint32_t half_bits = 1;
  while ((1u << (2 * half_bits)) < n) ++half_bits;
  const uint32_t v = feistel_perm((uint32_t)k, n, half_bits, seed);
  const int row = row0 + (int)(v / (uint32_t)win_w), col = col0 + (int)(v % (uint32_t)win_w);
  const float* c2w = poses + (int64_t)img * 12;  // [3][4]
  // the ray of that pixel: ray_util.py:75,78 exactly as get_rays_kernel forms it
  const float d0 = __fdiv_rn(__fsub_rn((float)col, cx), fx);
  const float d1 = __fdiv_rn(-__fsub_rn((float)row, cy), fy);
  const float d2 = -1.f;
  float dir[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* m = c2w + c * 4;
    dir[c] = __fadd_rn(__fadd_rn(__fmul_rn(d0, __ldg(m)), __fmul_rn(d1, __ldg(m + 1))), __fmul_rn(d2, __ldg(m + 2)));
  }
  float* dst = rays + k * 11;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dst[c] = __ldg(c2w + c * 4 + 3);
    dst[3 + c] = dir[c];
  }
  dst[6] = near;
  dst[7] = far;
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dir[0], dir[0]), __fmul_rn(dir[1], dir[1])), __fmul_rn(dir[2], dir[2])));
  dst[8] = __fdiv_rn(dir[0], nrm);   // run_nerf_helpers.py:350
  dst[9] = __fdiv_rn(dir[1], nrm);
  dst[10] = __fdiv_rn(dir[2], nrm);
  const float* px = images + (((int64_t)img * H + row) * W + col) * C;
#pragma unroll
  for (int c = 0; c < 3; ++c) target[k * 3 + c] = __ldg(px + c);
  if (pix != nullptr) {
    pix[k * 2] = row;
    pix[k * 2 + 1] = col;
  }
}

}  // namespace hn

extern "C" {

int hn_ndc_rays(int H, int W, double focal, double near, const float* rays_o, int64_t o_stride, const float* rays_d,
                int64_t d_stride, int64_t R, float* out_o, float* out_d, void* stream) {
  HN_REQUIRE(R >= 0, "hn_ndc_rays: negative R");
  HN_REQUIRE(H >= 1 && W >= 1 && focal != 0.0, "hn_ndc_rays: bad camera");
  if (R == 0) return 0;
  HN_REQUIRE(rays_o && rays_d && out_o && out_d, "hn_ndc_rays: null pointer");
  const float sx = (float)(-1.0 / ((double)W / (2.0 * focal))), sy = (float)(-1.0 / ((double)H / (2.0 * focal)));
  hn::ndc_rays_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, o_stride, rays_d, d_stride, R, (float)near, (float)(2.0 * near), sx, sy, out_o, out_d);
  return hn::check_launch("ndc_rays_kernel");
}

int hn_sample_rays(const float* images, int channels, const float* poses, int H, int W, float fx, float fy, float cx,
                   float cy, float near, float far, const int32_t* step, int64_t n_rand, float* rays, float* target,
                   int32_t* pix, void* stream) {
  HN_REQUIRE(n_rand >= 0, "hn_sample_rays: negative n_rand");
  HN_REQUIRE(H >= 1 && W >= 1 && (int64_t)H * W < (int64_t)1 << 30, "hn_sample_rays: bad image size");
  HN_REQUIRE(channels >= 3, "hn_sample_rays: images need at least 3 channels");
  if (n_rand == 0) return 0;
  HN_REQUIRE(images && poses && step && rays && target, "hn_sample_rays: null pointer");
  hn::sample_rays_kernel<<<(unsigned)((n_rand + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      images, channels, poses, H, W, fx, fy, cx, cy, near, far, step, n_rand, rays, target, pix);
  return hn::check_launch("sample_rays_kernel");
}

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
