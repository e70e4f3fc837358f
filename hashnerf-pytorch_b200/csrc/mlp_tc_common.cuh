// mlp_tc_common.cuh -- pieces shared by the tcgen05 MLP kernels (mlp_tc.cu: forward and the two-kernel 3xTF32
// backward; mlp_tc_bwd.cu: the fused bf16x2 backward).
#pragma once
#include "mlp_common.cuh"
#include "tc05.cuh"

namespace hn {
namespace tc {

constexpr int kTile = 128;  // points per tile = threads per tile context = TMEM lanes

// everything a layer boundary needs: stores visible -> the tile's 128 threads arrived -> one of them issues ->
// all wait.  `sync_id` names the hardware barrier of this tile context (0 = the CTA-wide barrier when the CTA
// runs a single context); `leader` is true for the context's issuing thread.
__device__ __forceinline__ void ctx_sync(int sync_id) {
  if (sync_id == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(sync_id), "r"(kTile) : "memory");
}

// Per-point inputs of one tile, loaded one tile ahead so that the global-load latency hides behind the layer
// chain of the current tile: the 32 hash features, the 16 SH coefficients of the point's ray, the keep flag.
struct TileInputs {
  float e[2][16];
  float v[16];
  uint8_t keep;
};

__device__ __forceinline__ void load_tile_inputs(TileInputs& in, int64_t p, int64_t N, const float* __restrict__ enc,
                                                 int64_t enc_stride, const float* __restrict__ views,
                                                 int64_t views_stride, int64_t pts_per_view,
                                                 const uint8_t* __restrict__ keep, int aligned) {
  const bool valid = p < N;
  const int64_t q = valid ? p : 0;
  const float* erow = enc + q * enc_stride;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (aligned) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(erow + 16 * h) + k);
        in.e[h][4 * k] = f.x;
        in.e[h][4 * k + 1] = f.y;
        in.e[h][4 * k + 2] = f.z;
        in.e[h][4 * k + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) in.e[h][i] = __ldg(erow + 16 * h + i);
    }
  }
  const float* vrow = views + (q / pts_per_view) * views_stride;
  if ((views_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(views) & 15) == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(vrow) + k);
      in.v[4 * k] = f.x;
      in.v[4 * k + 1] = f.y;
      in.v[4 * k + 2] = f.z;
      in.v[4 * k + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) in.v[i] = __ldg(vrow + i);
  }
  in.keep = (keep != nullptr) ? __ldg(keep + q) : (uint8_t)1;
  if (!valid) {
#pragma unroll
    for (int i = 0; i < 16; ++i) in.e[0][i] = in.e[1][i] = in.v[i] = 0.f;
  }
}

}  // namespace tc
}  // namespace hn
