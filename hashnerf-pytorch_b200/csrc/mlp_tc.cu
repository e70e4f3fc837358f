// mlp_tc.cu -- NeRFSmall on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators and
// activations in TMEM), 3xTF32 split so that the result keeps fp32-level accuracy.
//
// Same contract as mlp.cu (reference models.py:151-174 + the expand/cat/mask of run_nerf_helpers.py:219-225).
//
// A CTA of 128 threads owns a tile of 128 points: thread t <-> point t <-> TMEM lane t.  Every layer is
//   D[128 x N] (TMEM) = A[128 x K] (TMEM) . W[N x K]^T (shared memory, canonical no-swizzle K-major image)
// issued by one thread as three chains of K/8 MMAs (A_lo.W_hi + A_hi.W_lo + A_hi.W_hi), committed to an
// mbarrier.  The epilogue is row-private: each thread reads its row of D with tcgen05.ld, applies ReLU /
// slicing / concatenation in registers, splits into tf32 hi + lo and writes the next layer's A operand straight
// back into TMEM with tcgen05.st -- activations never touch shared or global memory.  Shared memory holds only
// the weights (hi and lo images, 76 KB), so two CTAs (= two tiles in flight) fit per SM; TMEM: 192 of the 256
// columns allocated per CTA (D 64 | A_hi 64 | A_lo 64).
#include "mlp_common.cuh"
#include "mlp_tc_common.cuh"

namespace hn {
namespace tc {

// canonical K-major weight images, in floats: [N][K] each
constexpr int oW0 = 0;                 // 64 x 32
constexpr int oW1 = oW0 + 64 * 32;     // 16 x 64
constexpr int oW2 = oW1 + 16 * 64;     // 64 x 32 (K padded 31 -> 32)
constexpr int oW3 = oW2 + 64 * 32;     // 64 x 64
constexpr int oW4 = oW3 + 64 * 64;     //  8 x 64 (N padded 3 -> 8)
constexpr int kImg = oW4 + 8 * 64;     // 9728 floats per image (hi, then lo)
constexpr uint32_t kColD = 0, kColAhi = 64, kColAlo = 128, kTmemCols = 256;

// kTile threads; shapes are compile-time so that the loop unrolls and eight weight loads are in flight per thread
// (one load per trip left every CTA's prologue at ~8 us of L2 round trips: 76 dependent trips)
template <int ROWS, int COLS, int ROWS_PAD, int K>
__device__ __forceinline__ void stage_matrix(const float* __restrict__ w, float* __restrict__ hi_img,
                                             float* __restrict__ lo_img) {
  static_assert((ROWS_PAD * K) % kTile == 0, "whole trips");
#pragma unroll 8
  for (int j = 0; j < (ROWS_PAD * K) / kTile; ++j) {
    const int i = j * kTile + (int)threadIdx.x;
    const int n = i / K, k = i % K;
    const float v = (n < ROWS && k < COLS) ? __ldg(w + n * COLS + k) : 0.f;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    hi_img[canon(n, k, K)] = __uint_as_float(hi);
    lo_img[canon(n, k, K)] = __uint_as_float(lo);
  }
}

// run-time shapes, any CTA size (the two-kernel backward's prologue)
__device__ __forceinline__ void stage_matrix_dyn(const float* __restrict__ w, int rows, int cols, int rows_pad, int K,
                                                 float* __restrict__ hi_img, float* __restrict__ lo_img) {
  for (int i = threadIdx.x; i < rows_pad * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const float v = (n < rows && k < cols) ? __ldg(w + n * cols + k) : 0.f;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    hi_img[canon(n, k, K)] = __uint_as_float(hi);
    lo_img[canon(n, k, K)] = __uint_as_float(lo);
  }
}

__device__ __forceinline__ void stage_weights(const float* __restrict__ w, float* __restrict__ smem) {
  float* hi = smem;
  float* lo = smem + kImg;
  stage_matrix<64, 32, 64, 32>(w + kG0, hi + oW0, lo + oW0);
  stage_matrix<16, 64, 16, 64>(w + kG1, hi + oW1, lo + oW1);
  stage_matrix<64, 31, 64, 32>(w + kG2, hi + oW2, lo + oW2);
  stage_matrix<64, 64, 64, 64>(w + kG3, hi + oW3, lo + oW3);
  stage_matrix<3, 64, 8, 64>(w + kG4, hi + oW4, lo + oW4);
}

// One layer: D = A . W^T as 3 x (K/8) MMAs; small terms first.  Called by ONE thread: its instruction stream is
// on the critical path of every layer, so N and K are template parameters (descriptor constants fold, the loops
// unroll and a K step is one 64-bit add of 16 = 256 bytes >> 4 on the descriptor).
template <int N, int K>
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t w_hi_saddr, uint32_t w_lo_saddr) {
  constexpr uint32_t idesc = make_idesc(kTile, N);
  constexpr uint64_t kHiBits = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(((K / 4) * 128) >> 4) << 32) | (1ull << 46);
  const uint64_t d_hi = kHiBits | (uint64_t)((w_hi_saddr >> 4) & 0x3FFF);
  const uint64_t d_lo = kHiBits | (uint64_t)((w_lo_saddr >> 4) & 0x3FFF);
#pragma unroll
  for (int s = 0; s < K / 8; ++s) umma_ts(tmem + kColD, tmem + kColAlo + 8 * s, d_hi + 16 * s, idesc, s ? 1u : 0u);
#pragma unroll
  for (int s = 0; s < K / 8; ++s) umma_ts(tmem + kColD, tmem + kColAhi + 8 * s, d_lo + 16 * s, idesc, 1u);
#pragma unroll
  for (int s = 0; s < K / 8; ++s) umma_ts(tmem + kColD, tmem + kColAhi + 8 * s, d_hi + 16 * s, idesc, 1u);
}

// write 16 activations (columns c0 .. c0+15 of this thread's row) as the next layer's A operand
__device__ __forceinline__ void put16(uint32_t row_taddr, int c0, const float (&v)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split_tf32(v[i], hi[i], lo[i]);
  tmem_st16(row_taddr + kColAhi + c0, hi);
  tmem_st16(row_taddr + kColAlo + c0, lo);
}

template <int N, int K>
__device__ __forceinline__ void run_layer(uint32_t tmem, uint64_t* bar, uint32_t& phase, uint32_t w_hi, uint32_t w_lo,
                                          int sync_id, bool leader_warp) {
  wait_st();
  fence_before_sync();
  ctx_sync(sync_id);
  if (leader_warp) {  // warp-uniform branch, one elected lane issues: descriptors stay in uniform registers
    if (elect_one()) {
      fence_after_sync();
      issue_layer<N, K>(tmem, w_hi, w_lo);
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, phase);
  phase ^= 1u;
  fence_after_sync();
}

// ReLU epilogue of a 64-wide hidden layer: D -> relu -> next A operand; GATES: also returns the mask of strictly
// positive pre-activations (what autograd would remember; handed to the backward kernel so that its gates are
// the forward's, whatever precision it recomputes the activations in).
template <bool GATES>
__device__ __forceinline__ uint64_t relu_epilogue64(uint32_t row) {
  uint64_t mask = 0;
  float v[4][16];
  tmem_ld64(row + kColD, v);  // all four loads in flight, one wait
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (GATES) {  // four independent 4-bit chains per quarter instead of one 64-long chain of predicated ORs
      uint32_t part[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int i = 0; i < 16; ++i) part[i >> 2] |= (v[q][i] > 0.f) ? (1u << i) : 0u;
      mask |= (uint64_t)((part[0] | part[1]) | (part[2] | part[3])) << (16 * q);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[q][i] = fmaxf(v[q][i], 0.f);
    put16(row, 16 * q, v[q]);
  }
  return mask;
}

template <bool GATES>
__global__ void __launch_bounds__(kTile, 2)
mlp_tc_fwd_kernel(const float* __restrict__ enc, int64_t enc_stride, const float* __restrict__ views,
                  int64_t views_stride, int64_t pts_per_view, const float* __restrict__ weights,
                  const uint8_t* __restrict__ keep, int64_t N, float* __restrict__ out, uint32_t* __restrict__ gates,
                  int aligned) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);  // warp-uniform for the compiler too
  const bool leader_warp = (warp == 0);
  stage_weights(weights, smem);
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);  // this thread's lane, column 0
  const uint32_t s_hi = smem_u32(smem), s_lo = smem_u32(smem + kImg);
  uint32_t phase = 0;

  const int64_t n_tiles = (N + kTile - 1) / kTile;
  TileInputs cur;
  if ((int64_t)blockIdx.x < n_tiles)
    load_tile_inputs(cur, (int64_t)blockIdx.x * kTile + t, N, enc, enc_stride, views, views_stride, pts_per_view, keep,
                     aligned);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * kTile + t;
    const bool valid = p < N;
    // ---- layer 0 input: the 32 hash features of this point
    put16(row, 0, cur.e[0]);
    put16(row, 16, cur.e[1]);
    float vsh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) vsh[i] = cur.v[i];
    const uint8_t keep_cur = cur.keep;
    // next tile's inputs: in flight while this tile runs its five layers
    if (tile + gridDim.x < n_tiles)
      load_tile_inputs(cur, (tile + gridDim.x) * kTile + t, N, enc, enc_stride, views, views_stride, pts_per_view, keep,
                       aligned);
    run_layer<64, 32>(tmem, &bar, phase, s_hi + oW0 * 4, s_lo + oW0 * 4, 0, leader_warp);  // h1 pre-activation
    const uint64_t m1 = relu_epilogue64<GATES>(row);
    fence_before_sync();  // D has been read: the next MMA may overwrite it after the barrier
    run_layer<16, 64>(tmem, &bar, phase, s_hi + oW1 * 4, s_lo + oW1 * 4, 0, leader_warp);  // h2 = [sigma | geo]
    float sigma;
    {
      float h2[16];
      tmem_ld16(row + kColD, h2);
      sigma = h2[0];
      // c = [views(16) | geo(15) | 0]
      float v[16];
      put16(row, 0, vsh);
#pragma unroll
      for (int i = 0; i < 15; ++i) v[i] = h2[1 + i];
      v[15] = 0.f;
      put16(row, 16, v);
    }
    fence_before_sync();
    run_layer<64, 32>(tmem, &bar, phase, s_hi + oW2 * 4, s_lo + oW2 * 4, 0, leader_warp);  // h3
    const uint64_t m3 = relu_epilogue64<GATES>(row);
    fence_before_sync();
    run_layer<64, 64>(tmem, &bar, phase, s_hi + oW3 * 4, s_lo + oW3 * 4, 0, leader_warp);  // h4
    const uint64_t m4 = relu_epilogue64<GATES>(row);
    fence_before_sync();
    run_layer<8, 64>(tmem, &bar, phase, s_hi + oW4 * 4, s_lo + oW4 * 4, 0, leader_warp);  // rgb (N padded to 8)
    {
      float rgb[8];
      tmem_ld8(row + kColD, rgb);
      if (valid) {
        const float s = (keep_cur == 0) ? 0.f : sigma;  // run_nerf_helpers.py:225
        reinterpret_cast<float4*>(out)[p] = make_float4(rgb[0], rgb[1], rgb[2], s);
        if (GATES) {
          uint2* g = reinterpret_cast<uint2*>(gates + p * 6);
          g[0] = make_uint2((uint32_t)m1, (uint32_t)(m1 >> 32));
          g[1] = make_uint2((uint32_t)m3, (uint32_t)(m3 >> 32));
          g[2] = make_uint2((uint32_t)m4, (uint32_t)(m4 >> 32));
        }
      }
    }
    fence_before_sync();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

constexpr size_t kFwdSmemBytes = (size_t)2 * kImg * sizeof(float);  // 77,824

// ================================================================================================
// backward, kernel 1: recompute + dX chain, all on tcgen05, activations/deltas resident in TMEM.
// Everything the weight gradients need is written feature-major to the workspace:
//   per 128-point tile two [472][64] fp32 half blocks (points 0..63 | 64..127); row r, point p of a half at
//   r * 64 + p  (a warp still stores 128 contiguous bytes; a half tile's rows of one operand are contiguous)
// ================================================================================================
constexpr int kWsStride = 64;  // samples per workspace row: a tile is stored as two 64-point half blocks
constexpr int rH1 = 0, rC = 64, rH3 = 96, rH4 = 160, rDz1 = 224, rDh2 = 288, rDz3 = 304, rDz4 = 368, rIn = 432,
              rDrgb = 464, kWsRowsTc = 472;
// transposed weight images for the dX chain (canonical K-major [N][K]), appended after the forward images
constexpr int oT4 = kImg;                // 64 x 8   : (n = hidden k, kk = c)      = W4[c][k]
constexpr int oT3 = oT4 + 64 * 8;        // 64 x 64  : (n = k, kk = j)             = W3[j][k]
constexpr int oT2 = oT3 + 64 * 64;       // 16 x 64  : (n = geo index, kk = j)     = W2[j][16 + n]   (n = 15: zero)
constexpr int oT1 = oT2 + 16 * 64;       // 64 x 16  : (n = k, kk = j)             = W1[j][k]
constexpr int oT0 = oT1 + 64 * 16;       // 32 x 64  : (n = k, kk = j)             = W0[j][k]
constexpr int kImgBwd = oT0 + 32 * 64;   // 18432 floats per image
constexpr size_t kDeltaSmemBytes = (size_t)2 * kImgBwd * sizeof(float);  // 147,456

// image(n, kk) = W[kk][col0 + n] for n < n_valid (W is [rows_w][cols_w] row-major), zero otherwise
__device__ __forceinline__ void stage_transposed(const float* __restrict__ w, int cols_w, int rows_w, int col0,
                                                 int n_valid, int n_rows, int K, float* __restrict__ hi_img,
                                                 float* __restrict__ lo_img) {
  for (int i = threadIdx.x; i < n_rows * K; i += blockDim.x) {
    const int n = i / K, kk = i % K;
    const float v = (n < n_valid && kk < rows_w) ? __ldg(w + kk * cols_w + col0 + n) : 0.f;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    hi_img[canon(n, kk, K)] = __uint_as_float(hi);
    lo_img[canon(n, kk, K)] = __uint_as_float(lo);
  }
}

template <bool RELU>
__device__ __forceinline__ uint64_t epilogue64(uint32_t row, float* __restrict__ ws_col, uint64_t gate, bool gated) {
  // reads this thread's 64 accumulator columns, applies ReLU (returning the positive mask) or the gate,
  // stores to the workspace column (stride 128) and writes the next A operand
  uint64_t mask = 0;
  float v[4][16];
  tmem_ld64(row + kColD, v);  // all four loads in flight, one wait
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c0 = 16 * q;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (RELU) {
        if (v[q][i] > 0.f) mask |= (1ull << (c0 + i));
        v[q][i] = fmaxf(v[q][i], 0.f);
      }
      if (gated) v[q][i] = ((gate >> (c0 + i)) & 1ull) ? v[q][i] : 0.f;
      ws_col[(c0 + i) * kWsStride] = v[q][i];
    }
    put16(row, c0, v[q]);
  }
  return mask;
}

// Two tile contexts of 128 threads share one copy of the weight images (144 KB) and interleave on the SM:
// while one context waits for its MMA chain the other runs its epilogue.
__global__ void __launch_bounds__(2 * kTile, 1)
mlp_tc_bwd_delta_kernel(const float* __restrict__ enc, int64_t enc_stride, const float* __restrict__ views,
                        int64_t views_stride, int64_t pts_per_view, const float* __restrict__ weights,
                        const uint8_t* __restrict__ keep, const float* __restrict__ dout, int64_t N,
                        float* __restrict__ d_enc, float* __restrict__ ws, int aligned) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int ctx = threadIdx.x / kTile, t = threadIdx.x % kTile, warp = t >> 5;
  const int sync_id = 1 + ctx;
  const bool leader = (warp == 0);  // warp-level: run_layer elects one lane
  uint64_t* bar_p = &bars[ctx];
  {
    float* hi = smem;
    float* lo = smem + kImgBwd;
    stage_matrix_dyn(weights + kG0, 64, 32, 64, 32, hi + oW0, lo + oW0);
    stage_matrix_dyn(weights + kG1, 16, 64, 16, 64, hi + oW1, lo + oW1);
    stage_matrix_dyn(weights + kG2, 64, 31, 64, 32, hi + oW2, lo + oW2);
    stage_matrix_dyn(weights + kG3, 64, 64, 64, 64, hi + oW3, lo + oW3);
    stage_transposed(weights + kG4, 64, 3, 0, 64, 64, 8, hi + oT4, lo + oT4);
    stage_transposed(weights + kG3, 64, 64, 0, 64, 64, 64, hi + oT3, lo + oT3);
    stage_transposed(weights + kG2, 31, 64, 16, 15, 16, 64, hi + oT2, lo + oT2);
    stage_transposed(weights + kG1, 64, 16, 0, 64, 64, 16, hi + oT1, lo + oT1);
    stage_transposed(weights + kG0, 32, 64, 0, 32, 32, 64, hi + oT0, lo + oT0);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 2 * kTmemCols);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot + (uint32_t)ctx * kTmemCols;
  const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t s_hi = smem_u32(smem), s_lo = smem_u32(smem + kImgBwd);
  uint32_t phase = 0;

  const int64_t n_tiles = (N + kTile - 1) / kTile;
  const int64_t tile_step = (int64_t)gridDim.x * 2;
  TileInputs cur;
  float4 go_cur = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    const int64_t tile0 = (int64_t)blockIdx.x * 2 + ctx;
    if (tile0 < n_tiles) {
      const int64_t p0 = tile0 * kTile + t;
      load_tile_inputs(cur, p0, N, enc, enc_stride, views, views_stride, pts_per_view, keep, aligned);
      if (p0 < N) go_cur = __ldg(reinterpret_cast<const float4*>(dout) + p0);
    }
  }
  for (int64_t tile = (int64_t)blockIdx.x * 2 + ctx; tile < n_tiles; tile += tile_step) {
    const int64_t p = tile * kTile + t;
    const bool valid = p < N;
    // this point's workspace column: half block t / 64, column t % 64 (a warp still writes 128 contiguous bytes)
    float* g = ws + tile * (int64_t)(kWsRowsTc * kTile) + (t >> 6) * (kWsRowsTc * kWsStride) + (t & 63);
    // ---- inputs (loaded one tile ahead)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 16; ++i) g[(rIn + 16 * h + i) * kWsStride] = cur.e[h][i];
      put16(row, 16 * h, cur.e[h]);
    }
    float vsh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) vsh[i] = cur.v[i];
    const float4 go = go_cur;
    const float dsigma = (valid && cur.keep != 0) ? go.w : 0.f;
    g[(rDrgb + 0) * kWsStride] = go.x;
    g[(rDrgb + 1) * kWsStride] = go.y;
    g[(rDrgb + 2) * kWsStride] = go.z;
#pragma unroll
    for (int i = 3; i < 8; ++i) g[(rDrgb + i) * kWsStride] = 0.f;
    // next tile's inputs: in flight while this tile runs its nine layers
    go_cur = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tile + tile_step < n_tiles) {
      const int64_t pn = (tile + tile_step) * kTile + t;
      load_tile_inputs(cur, pn, N, enc, enc_stride, views, views_stride, pts_per_view, keep, aligned);
      if (pn < N) go_cur = __ldg(reinterpret_cast<const float4*>(dout) + pn);
    }

    // ---- forward recompute
    run_layer<64, 32>(tmem, bar_p, phase, s_hi + oW0 * 4, s_lo + oW0 * 4, sync_id, leader);
    const uint64_t m1 = epilogue64<true>(row, g + rH1 * kWsStride, 0, false);
    fence_before_sync();
    run_layer<16, 64>(tmem, bar_p, phase, s_hi + oW1 * 4, s_lo + oW1 * 4, sync_id, leader);
    {
      float h2[16];
      tmem_ld16(row + kColD, h2);
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) g[(rC + i) * kWsStride] = vsh[i];
      put16(row, 0, vsh);
#pragma unroll
      for (int i = 0; i < 15; ++i) v[i] = h2[1 + i];
      v[15] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) g[(rC + 16 + i) * kWsStride] = v[i];
      put16(row, 16, v);
    }
    fence_before_sync();
    run_layer<64, 32>(tmem, bar_p, phase, s_hi + oW2 * 4, s_lo + oW2 * 4, sync_id, leader);
    const uint64_t m3 = epilogue64<true>(row, g + rH3 * kWsStride, 0, false);
    fence_before_sync();
    run_layer<64, 64>(tmem, bar_p, phase, s_hi + oW3 * 4, s_lo + oW3 * 4, sync_id, leader);
    uint64_t m4 = 0;
    {  // h4: only its values (for dW4) and its mask are needed
      float v[4][16];
      tmem_ld64(row + kColD, v);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (v[q][i] > 0.f) m4 |= (1ull << (16 * q + i));
          g[(rH4 + 16 * q + i) * kWsStride] = fmaxf(v[q][i], 0.f);
        }
    }
    // ---- backward chain
    {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
      v[0] = go.x;
      v[1] = go.y;
      v[2] = go.z;
      put16(row, 0, v);  // A = drgb, K = 8 (columns 3..7 zero)
    }
    fence_before_sync();
    run_layer<64, 8>(tmem, bar_p, phase, s_hi + oT4 * 4, s_lo + oT4 * 4, sync_id, leader);    // dh4 = drgb . W4
    epilogue64<false>(row, g + rDz4 * kWsStride, m4, true);                      // dz4 = dh4 . [h4 > 0]
    fence_before_sync();
    run_layer<64, 64>(tmem, bar_p, phase, s_hi + oT3 * 4, s_lo + oT3 * 4, sync_id, leader);   // dh3 = dz4 . W3
    epilogue64<false>(row, g + rDz3 * kWsStride, m3, true);                      // dz3
    fence_before_sync();
    run_layer<16, 64>(tmem, bar_p, phase, s_hi + oT2 * 4, s_lo + oT2 * 4, sync_id, leader);   // dgeo = (dz3 . W2)[16:31]
    {
      float dg[16], v[16];
      tmem_ld16(row + kColD, dg);
      v[0] = dsigma;
#pragma unroll
      for (int i = 0; i < 15; ++i) v[1 + i] = dg[i];
#pragma unroll
      for (int i = 0; i < 16; ++i) g[(rDh2 + i) * kWsStride] = v[i];
      put16(row, 0, v);  // A = dh2, K = 16
    }
    fence_before_sync();
    run_layer<64, 16>(tmem, bar_p, phase, s_hi + oT1 * 4, s_lo + oT1 * 4, sync_id, leader);   // dh1 = dh2 . W1
    epilogue64<false>(row, g + rDz1 * kWsStride, m1, true);                      // dz1
    fence_before_sync();
    run_layer<32, 64>(tmem, bar_p, phase, s_hi + oT0 * 4, s_lo + oT0 * 4, sync_id, leader);   // d_enc = dz1 . W0
#pragma unroll
    for (int c0 = 0; c0 < kIn; c0 += 16) {
      float v[16];
      tmem_ld16(row + kColD + c0, v);
      if (valid) {
        float* drow = d_enc + p * kIn + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(drow)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    fence_before_sync();
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_slot, 2 * kTmemCols);
}

// ================================================================================================
// backward, kernel 2: weight gradients on tcgen05.  dW[j][k] = sum_p D[p][j] * A[p][k] is an MMA with the
// POINTS as the K dimension; the feature-major workspace rows are exactly K-major operands.  M = 64 feature
// rows, N = the other tensor's features, K = 8 points per instruction.  Accumulators stay in TMEM for the
// whole kernel (152 columns) and are flushed once with atomics.  A CTA streams 64-point half tiles:
// load rows -> split hi/lo -> 128-byte-swizzled K-major image in shared memory -> 3 x 8 MMAs per matrix.
// ================================================================================================
constexpr int kHalf = 64;
constexpr uint32_t cW0 = 0, cW1 = 32, cW2 = 48, cW3 = 80, cW4 = 144;  // accumulator columns
constexpr size_t kWeightBufBytes = (size_t)4 * 64 * kHalf * sizeof(float);  // (M_hi | M_lo | N_hi | N_lo) = 64 KB
constexpr size_t kWeightAlignPad = 1024;  // slack so that the staging area can start on a 1024-byte boundary

// Staging of one operand pair for a 64-point half tile.  The rows of one operand are consecutive workspace rows of
// 64 floats, i.e. ONE contiguous range, so thread t's slot s simply takes float4 number s*256 + t of the pair
// (the first 64*16 belong to the M-side rows, the rest to the N-side rows): a quarter warp reads one full
// 128-byte line.  In shared memory an operand is two 128-byte-swizzled K-major blocks (points 0..31 and 32..63
// of the half tile); the 8 lanes of a quarter warp hold the 8 chunks of one row and the XOR with the row index
// spreads them over all 32 banks (conflict-free STS.128).
constexpr int kWThreads = 256;  // threads of the weight-gradient kernel
constexpr int kSlots = 8;       // (64 + 64 rows) * 16 float4 / 256 threads

struct PairDesc {
  int m_row, n_row, n_rows;
  uint32_t col;
};

__device__ __forceinline__ void prefetch_pair(const float* __restrict__ base, const PairDesc& pr, float4 (&reg)[kSlots]) {
  const int total = (64 + pr.n_rows) * (kHalf / 4);
  const float4* m_src = reinterpret_cast<const float4*>(base + (int64_t)pr.m_row * kWsStride);
  const float4* n_src = reinterpret_cast<const float4*>(base + (int64_t)pr.n_row * kWsStride) - 64 * (kHalf / 4);
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int idx = s * kWThreads + threadIdx.x;
    reg[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < total) reg[s] = __ldg((idx >= 64 * (kHalf / 4) ? n_src : m_src) + idx);
  }
}

__device__ __forceinline__ void store_pair(const PairDesc& pr, const float4 (&reg)[kSlots], float* __restrict__ Mhi,
                                           float* __restrict__ Mlo, float* __restrict__ Nhi, float* __restrict__ Nlo) {
  const int total = (64 + pr.n_rows) * (kHalf / 4);
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int idx = s * kWThreads + threadIdx.x;
    if (idx < total) {
      const bool is_n = idx >= 64 * (kHalf / 4);
      const int i = is_n ? idx - 64 * (kHalf / 4) : idx;
      const int row = i >> 4, kb = (i >> 3) & 1, j = i & 7;     // 16 float4 per row: 2 blocks of 8 chunks
      const int rows = is_n ? pr.n_rows : 64;
      const int off = kb * (rows * 32) + (row >> 3) * 256 + (row & 7) * 32 + ((j ^ (row & 7)) << 2);
      uint32_t h[4], l[4];
      split_tf32(reg[s].x, h[0], l[0]);
      split_tf32(reg[s].y, h[1], l[1]);
      split_tf32(reg[s].z, h[2], l[2]);
      split_tf32(reg[s].w, h[3], l[3]);
      *reinterpret_cast<uint4*>((is_n ? Nhi : Mhi) + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>((is_n ? Nlo : Mlo) + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

// 256 threads.  NBUF = 2: one CTA per SM, two staging buffers -- while the tensor core consumes pair q from
// buffer q & 1 the threads split and store pair q + 1 into the other buffer and the global loads of pair q + 2
// are in flight.  NBUF = 1: one staging buffer and two CTAs per SM -- the second CTA fills the bubbles of the
// first (its MMA wait, its barrier, its load latency) instead of a second buffer.
template <int NBUF>
__global__ void __launch_bounds__(kWThreads, NBUF == 1 ? 2 : 1)
mlp_tc_bwd_weight_kernel(int64_t N, const float* __restrict__ ws, float* __restrict__ dweights, int ablate) {
  extern __shared__ __align__(128) float smem_raw[];
  // swizzle atoms are addressed by absolute shared-memory address bits: align the staging area to 1024 bytes
  float* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  uint32_t phase[2] = {0u, 0u};
  bool used[2] = {false, false};  // an MMA group reading this staging buffer has been committed
  uint32_t fresh = 0x1f;          // bit i set: accumulator i has not been written yet (first MMA overwrites)

  // (M-side rows, N-side rows): dW0 = dz1^T.in, dW1^T = h1^T.dh2, dW2 = dz3^T.c, dW3 = dz4^T.h3, dW4^T = h4^T.drgb
  const PairDesc pairs[5] = {{rDz1, rIn, 32, cW0}, {rH1, rDh2, 16, cW1}, {rDz3, rC, 32, cW2}, {rDz4, rH3, 64, cW3},
                             {rH4, rDrgb, 8, cW4}};
  const int64_t n_half = ((N + kTile - 1) / kTile) * 2;
  const int64_t G = gridDim.x;
  auto base_of = [&](int64_t hh) { return ws + (hh >> 1) * (int64_t)(kWsRowsTc * kTile) + (hh & 1) * (kWsRowsTc * kWsStride); };
  // Two register sets: the set consumed at step k is refilled at once with the loads of step k + 2, so global
  // loads have two full pair-times (several microseconds) to land.  A loop iteration covers two half tiles
  // = 10 steps, which keeps the set <-> step mapping static.
  float4 ra[kSlots], rb[kSlots];
  int64_t h = blockIdx.x;
  if (h < n_half) {
    prefetch_pair(base_of(h), pairs[0], ra);
    prefetch_pair(base_of(h), pairs[1], rb);
  }
  // profiling ablations (hn_set_tuning "mlp_dw_ablate"): 1 = no global loads after the first two, 2 = no MMAs,
  // 4 = no split/store.  Results are wrong with any of them; they only apportion the kernel's time.
  int q = 0;
  for (; h < n_half; h += 2 * G) {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const int64_t hh = h + (k / 5) * G;
      if (hh >= n_half) break;
      const int i = k % 5;
      float4(&cur)[kSlots] = (k & 1) ? rb : ra;
      const PairDesc pr = pairs[i];
      const int b = (NBUF == 2) ? (q & 1) : 0;
      ++q;
      float* Mhi = smem + b * (4 * 64 * kHalf);
      float* Mlo = Mhi + 64 * kHalf;
      float* Nhi = Mlo + 64 * kHalf;
      float* Nlo = Nhi + 64 * kHalf;
      if (used[b]) {  // the MMAs issued two steps ago must have consumed this buffer
        mbar_wait(&bars[b], phase[b]);
        phase[b] ^= 1u;
      }
      if (!(ablate & 4)) store_pair(pr, cur, Mhi, Mlo, Nhi, Nlo);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (t == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc(64, pr.n_rows);
        const uint64_t m_hi = make_sdesc_sw128(smem_u32(Mhi), 1024), m_lo = make_sdesc_sw128(smem_u32(Mlo), 1024);
        const uint64_t n_hi = make_sdesc_sw128(smem_u32(Nhi), 1024), n_lo = make_sdesc_sw128(smem_u32(Nlo), 1024);
        // K step s: block s / 4 (rows * 128 bytes further), 32 bytes per step inside the swizzled row
        const uint64_t m_blk = (uint64_t)((64 * 128) >> 4), n_blk = (uint64_t)((pr.n_rows * 128) >> 4);
        const uint32_t first = ((fresh >> i) & 1u) ? 0u : 1u;
        const uint32_t d = tmem + pr.col;
        if (!(ablate & 2)) {
#pragma unroll
          for (int s = 0; s < kHalf / 8; ++s) {
            const uint64_t mo = (s >> 2) * m_blk + 2 * (s & 3), no = (s >> 2) * n_blk + 2 * (s & 3);
            umma_ss(d, m_lo + mo, n_hi + no, idesc, s ? 1u : first);
          }
#pragma unroll
          for (int s = 0; s < kHalf / 8; ++s) {
            const uint64_t mo = (s >> 2) * m_blk + 2 * (s & 3), no = (s >> 2) * n_blk + 2 * (s & 3);
            umma_ss(d, m_hi + mo, n_lo + no, idesc, 1u);
          }
#pragma unroll
          for (int s = 0; s < kHalf / 8; ++s) {
            const uint64_t mo = (s >> 2) * m_blk + 2 * (s & 3), no = (s >> 2) * n_blk + 2 * (s & 3);
            umma_ss(d, m_hi + mo, n_hi + no, idesc, 1u);
          }
        }
        umma_commit(&bars[b]);
      }
      fresh &= ~(1u << i);
      used[b] = true;
      // refill this register set with step k + 2
      const int k2 = k + 2;
      const int64_t h2 = (k2 < 10) ? h + (k2 / 5) * G : h + 2 * G;
      if (h2 < n_half && !(ablate & 1)) prefetch_pair(base_of(h2), pairs[k2 % 5], cur);
    }
  }
#pragma unroll
  for (int b = 0; b < 2; ++b)
    if (used[b]) {
      mbar_wait(&bars[b], phase[b]);
      phase[b] ^= 1u;
    }
  fence_after_sync();
  // ---- flush: M = 64 accumulators keep row m in lane (m % 16) + 32 * (m / 16)
  const bool any = blockIdx.x < n_half;
  const int m = warp * 16 + lane;  // valid for lane < 16, warps 0..3 (each warp reads its own 32-lane quadrant)
  const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  if (any && warp < 4) {
    float v[16];
    // dW0[j = m][k]   (64 x 32)
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
      tmem_ld16(row + cW0 + c0, v);
      if (lane < 16)
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(dweights + kG0 + m * 32 + c0 + i, v[i]);
    }
    // dW1^T[k = m][j] -> dW1[j][k]   (16 x 64)
    tmem_ld16(row + cW1, v);
    if (lane < 16)
#pragma unroll
      for (int i = 0; i < 16; ++i) atomicAdd(dweights + kG1 + i * 64 + m, v[i]);
    // dW2[j = m][c]   (64 x 31)
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
      tmem_ld16(row + cW2 + c0, v);
      if (lane < 16)
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < 31) atomicAdd(dweights + kG2 + m * 31 + c0 + i, v[i]);
    }
    // dW3[j = m][k]   (64 x 64)
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      tmem_ld16(row + cW3 + c0, v);
      if (lane < 16)
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(dweights + kG3 + m * 64 + c0 + i, v[i]);
    }
    // dW4^T[k = m][c] -> dW4[c][k]   (3 x 64)
    {
      float u[8];
      tmem_ld8(row + cW4, u);
      if (lane < 16)
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(dweights + kG4 + c * 64 + m, u[c]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace tc

int g_mlp_fwd_one_cta = 0;

int mlp_tc_fwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
               const float* weights, const uint8_t* keep, int64_t N, float* out, uint32_t* gates, int aligned,
               cudaStream_t stream) {
  static thread_local int done_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, "cudaGetDevice");
  if (done_dev != dev) {
    e = cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_tc_fwd_kernel)");
    e = cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_tc_fwd_kernel)");
    done_dev = dev;
  }
  const int64_t tiles = (N + tc::kTile - 1) / tc::kTile;
  const int64_t cap = (int64_t)sm_count() * 2;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  // profiling knob (hn_set_tuning "mlp_fwd_one_cta"): pad the dynamic shared memory so that only one CTA fits per SM
  const size_t smem_bytes = g_mlp_fwd_one_cta ? (size_t)160 * 1024 : tc::kFwdSmemBytes;
  if (gates != nullptr)
    tc::mlp_tc_fwd_kernel<true><<<grid, tc::kTile, smem_bytes, stream>>>(enc, enc_stride, views, views_stride, pts_per_view,
                                                                        weights, keep, N, out, gates, aligned);
  else
    tc::mlp_tc_fwd_kernel<false><<<grid, tc::kTile, smem_bytes, stream>>>(enc, enc_stride, views, views_stride, pts_per_view,
                                                                         weights, keep, N, out, gates, aligned);
  return check_launch("mlp_tc_fwd_kernel");
}

int g_mlp_dw_ablate = 0;
int g_mlp_dw_nbuf = 1;  // 1: single staging buffer, two CTAs per SM; 2: double buffer, one CTA per SM

int64_t mlp_tc_bwd_workspace_floats(int64_t N) {
  const int64_t tiles = (N + tc::kTile - 1) / tc::kTile;
  return tiles * tc::kWsRowsTc * tc::kTile;
}

int mlp_tc_bwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
               const float* weights, const uint8_t* keep, const float* dout, int64_t N, float* d_enc, float* dweights,
               float* workspace, int aligned, cudaStream_t stream) {
  static thread_local int done_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, "cudaGetDevice");
  if (done_dev != dev) {
    e = cudaFuncSetAttribute(tc::mlp_tc_bwd_delta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)tc::kDeltaSmemBytes);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_tc_bwd_delta_kernel)");
    e = cudaFuncSetAttribute(tc::mlp_tc_bwd_weight_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(tc::kWeightBufBytes + tc::kWeightAlignPad));
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_tc_bwd_weight_kernel<1>)");
    e = cudaFuncSetAttribute(tc::mlp_tc_bwd_weight_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(2 * tc::kWeightBufBytes + tc::kWeightAlignPad));
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_tc_bwd_weight_kernel<2>)");
    done_dev = dev;
  }
  const int64_t tiles = (N + tc::kTile - 1) / tc::kTile;
  {
    const int64_t cap = (int64_t)sm_count();
    const int64_t want = (tiles + 1) / 2;  // two tile contexts per CTA
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    tc::mlp_tc_bwd_delta_kernel<<<grid, 2 * tc::kTile, tc::kDeltaSmemBytes, stream>>>(
        enc, enc_stride, views, views_stride, pts_per_view, weights, keep, dout, N, d_enc, workspace, aligned);
    int rc = check_launch("mlp_tc_bwd_delta_kernel");
    if (rc) return rc;
  }
  {
    const int64_t halves = tiles * 2;
    const int nbuf = g_mlp_dw_nbuf == 2 ? 2 : 1;
    const int64_t cap = (int64_t)sm_count() * (nbuf == 1 ? 2 : 1);
    const unsigned grid = (unsigned)(halves < cap ? halves : cap);
    if (nbuf == 1)
      tc::mlp_tc_bwd_weight_kernel<1><<<grid, tc::kWThreads, tc::kWeightBufBytes + tc::kWeightAlignPad, stream>>>(N, workspace, dweights,
                                                                                           g_mlp_dw_ablate);
    else
      tc::mlp_tc_bwd_weight_kernel<2><<<grid, tc::kWThreads, 2 * tc::kWeightBufBytes + tc::kWeightAlignPad, stream>>>(
          N, workspace, dweights, g_mlp_dw_ablate);
    return check_launch("mlp_tc_bwd_weight_kernel");
  }
}

}  // namespace hn
