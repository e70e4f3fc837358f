// common.cuh -- shared helpers for libhashnerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hashnerf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhashnerf_b200 is written for sm_100a (Blackwell B200) only"
#endif

namespace hn {

// Records the message returned by hn_last_error_string() and passes the code through.
int fail(int code, const char* what);
int check_launch(const char* kernel_name);
int sm_count();

#define HN_REQUIRE(cond, msg) \
  do {                        \
    if (!(cond)) return ::hn::fail(HN_EINVAL, msg); \
  } while (0)

// embedding/hash_encoding.py:7 -- multipliers of the spatial hash (uint32 wrap keeps the low bits).
__device__ __forceinline__ uint32_t hash3(uint32_t ix, uint32_t iy, uint32_t iz, uint32_t mask) {
  return (ix ^ (iy * 2654435761u) ^ (iz * 805459861u)) & mask;
}

struct Box {
  float lo[3];
  float hi[3];
};

__device__ __forceinline__ Box load_box(const float* __restrict__ bbox) {
  Box b;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    b.lo[a] = __ldg(bbox + a);
    b.hi[a] = __ldg(bbox + 3 + a);
  }
  return b;
}

// torch.max(torch.min(x, hi), lo) with NaN propagation (hash_encoding.py:66,69).
__device__ __forceinline__ float clamp_box(float x, float lo, float hi) {
  return (x != x) ? x : fmaxf(fminf(x, hi), lo);
}

// One axis of get_voxel_vertices + the interpolation weight (hash_encoding.py:72-76, 143).
// Every operation is an individually rounded fp32 op in the reference's order:
// no FMA contraction, IEEE division.
struct AxisCell {
  int idx;     // bottom-left voxel index
  float vmin;  // idx*g + lo
  float vmax;  // vmin + g
  float w;     // (x - vmin) / (vmax - vmin), x UNCLAMPED
};

__device__ __forceinline__ AxisCell axis_cell(float x, float xc, float lo, float hi, float res) {
  AxisCell c;
  const float g = __fdiv_rn(__fsub_rn(hi, lo), res);
  c.idx = (int)floorf(__fdiv_rn(__fsub_rn(xc, lo), g));
  c.vmin = __fadd_rn(__fmul_rn((float)c.idx, g), lo);
  c.vmax = __fadd_rn(c.vmin, g);
  c.w = __fdiv_rn(__fsub_rn(x, c.vmin), __fsub_rn(c.vmax, c.vmin));
  return c;
}

// a*(1-w) + b*w as the reference evaluates it: sub, mul, mul, add (hash_encoding.py:149-161).
__device__ __forceinline__ float lerp_ref(float a, float b, float w, float one_minus_w) {
  return __fadd_rn(__fmul_rn(a, one_minus_w), __fmul_rn(b, w));
}

}  // namespace hn
