// mlp_common.cuh -- geometry of the instantiated NeRFSmall and the packed weight layout shared by the FFMA and
// tcgen05 implementations (reference models.py:96-174 as built at run_nerf_helpers.py:79-84).
#pragma once
#include "common.cuh"

namespace hn {

constexpr int kNT = 128;  // threads per CTA == points per tile
constexpr int kIn = 32, kViews = 16, kHid = 64, kGeo = 15, kH2 = 16, kCin = 31, kCinPad = 32;
// shared-memory weight image (W2 rows padded 31 -> 32 so every row is float4-aligned)
constexpr int kW0 = 0;                     // [64][32]
constexpr int kW1 = kW0 + kHid * kIn;      // [16][64]
constexpr int kW2 = kW1 + kH2 * kHid;      // [64][32] (padded)
constexpr int kW3 = kW2 + kHid * kCinPad;  // [64][64]
constexpr int kW4 = kW3 + kHid * kHid;     // [3][64]
constexpr int kWTotal = kW4 + 3 * kHid;    // 9408 floats
// transposed image used by the backward pass: Wt[k][j] = W[j][k]
constexpr int kT0 = kWTotal;               // [32][64]  (W0^T)
constexpr int kT1 = kT0 + kIn * kHid;      // [64][16]  (W1^T)
constexpr int kT2 = kT1 + kHid * kH2;      // [32][64]  (W2^T, padded row 31 = 0)
constexpr int kT3 = kT2 + kCinPad * kHid;  // [64][64]  (W3^T)
constexpr int kWBoth = kT3 + kHid * kHid;  // 18624 floats
// packed global layout (nn.Linear.weight order, unpadded)
constexpr int kG0 = 0, kG1 = kG0 + 2048, kG2 = kG1 + 1024, kG3 = kG2 + 64 * 31, kG4 = kG3 + 4096;
static_assert(kG4 + 192 == HN_MLP_PARAMS, "packed weight count");

__device__ __forceinline__ float packed_weight(const float* __restrict__ w, int i) {
  // i indexes the padded shared image [kW0, kWTotal)
  if (i < kW1) return __ldg(w + kG0 + i);
  if (i < kW2) return __ldg(w + kG1 + (i - kW1));
  if (i < kW3) {
    const int r = (i - kW2) >> 5, c = (i - kW2) & 31;
    return (c < kCin) ? __ldg(w + kG2 + r * kCin + c) : 0.f;
  }
  if (i < kW4) return __ldg(w + kG3 + (i - kW3));
  return __ldg(w + kG4 + (i - kW4));
}

}  // namespace hn
