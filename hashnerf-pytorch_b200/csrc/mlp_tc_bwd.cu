// mlp_tc_bwd.cu -- NeRFSmall backward as ONE tcgen05 kernel: forward recompute, the dX chain AND the weight
// gradients, with nothing but d_enc leaving the SM (autograd of reference models.py:151-174 plus the
// expand/cat/mask of run_nerf_helpers.py:219-225).
//
// What changed against the two-kernel backward in mlp_tc.cu: that one wrote every activation and delta of a tile
// (472 rows x 128 points fp32 = 241 KB) to an HBM workspace and a second kernel read it back to form
// dW = sum_p delta[p]^T act[p] -- 6.1 GB of DRAM traffic per 1.57 M samples around 0.43 GB of real input/output.
// Here the thread that owns a point packs its activations / deltas once, as bf16 (hi, lo) pairs, and the same
// packed words feed (a) the next layer's A operand in TMEM and (b) a shared-memory line [point][64 features]
// that the tensor core reads as an MN-major operand: the contraction index of dW is the POINT, which is the slow
// index of that layout, so no transpose is ever materialised.
//
// Arithmetic: every operand x is split as x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (|x - hi - lo| <=
// 2^-18 |x|) and every product is evaluated as lo.hi + hi.lo + hi.hi on kind::f16 MMAs with fp32 accumulation:
// relative error about 1e-5 per product, against the 1e-4 bar for gradients (the forward kernel keeps 3xTF32 for
// its 1e-5 bar).  bf16 pairs halve the TMEM and shared-memory footprint of an operand against tf32 pairs, which
// is what lets two tile contexts, their weight images, their operand lines and all ten dW accumulators live on
// one SM:
//   TMEM   (512 columns): context c: D 64 | A_hi 32 | A_lo 32 at 128 c; dW accumulators of context c: 80 columns at
//          256 + 80 c, two M = 64 accumulators sharing a column range on interleaved lane halves
//          (dW3 | dW0, dW2 on lanes +16; dW1^T | dW4^T on lanes +16)
//   shared (209 KB): bf16 weight images 72 KB (built once by mlp_bwd_prep_kernel, fetched with one TMA bulk copy)
//          + per context X (M-side line buffer, 32 KB), Y (N-side, 32 KB), Z (drgb, 4 KB)
// A context is 128 threads = 128 points = 128 TMEM lanes; the two contexts of a CTA interleave on the SM (one runs
// its epilogue while the other's MMAs execute).  Buffer reuse needs no extra barriers: tcgen05.mma operations of a
// CTA execute in issue order and every step ends with one tcgen05.commit, so when a step's mbarrier fires all
// operand reads of the dW MMAs issued with it are over.
#include "mlp_tc_common.cuh"

namespace hn {
extern int g_mlp_dw_ablate;  // mlp_tc.cu (profiling only)

namespace tc {

// ---- bf16 canonical K-major weight images [N][K], in elements; hi image, then lo image -----------------------
constexpr int bW0 = 0;                 // 64 x 32  W0
constexpr int bW1 = bW0 + 64 * 32;     // 16 x 64  W1
constexpr int bW2 = bW1 + 16 * 64;     // 64 x 32  W2 (K padded 31 -> 32)
constexpr int bW3 = bW2 + 64 * 32;     // 64 x 64  W3
constexpr int bT4 = bW3 + 64 * 64;     // 64 x 16  (n = k, kk = c)     = W4[c][k], c < 3
constexpr int bT3 = bT4 + 64 * 16;     // 64 x 64  (n = k, kk = j)     = W3[j][k]
constexpr int bT2 = bT3 + 64 * 64;     // 16 x 64  (n = geo, kk = j)   = W2[j][16 + n], n < 15
constexpr int bT1 = bT2 + 16 * 64;     // 64 x 16  (n = k, kk = j)     = W1[j][k]
constexpr int bT0 = bT1 + 64 * 16;     // 32 x 64  (n = k, kk = j)     = W0[j][k]
constexpr int kImg16 = bT0 + 32 * 64;  // 18432 elements per image
constexpr uint32_t kImgBytes = (uint32_t)kImg16 * 2u * 2u;  // 73,728
constexpr int kGTotal = kG4 + 3 * 64;                       // 9344 weight-gradient floats
constexpr int64_t kGTotalBytes = (int64_t)kGTotal * 4;
static_assert(kImgBytes % 1024 == 0, "operand buffers behind the images must stay 1024-byte aligned");

__device__ __forceinline__ float image_value(const float* __restrict__ w, int mat, int n, int k) {
  switch (mat) {
    case 0: return __ldg(w + kG0 + n * 32 + k);
    case 1: return __ldg(w + kG1 + n * 64 + k);
    case 2: return k < 31 ? __ldg(w + kG2 + n * 31 + k) : 0.f;
    case 3: return __ldg(w + kG3 + n * 64 + k);
    case 4: return k < 3 ? __ldg(w + kG4 + k * 64 + n) : 0.f;
    case 5: return __ldg(w + kG3 + k * 64 + n);
    case 6: return n < 15 ? __ldg(w + kG2 + k * 31 + 16 + n) : 0.f;
    case 7: return __ldg(w + kG1 + k * 64 + n);
    default: return __ldg(w + kG0 + k * 32 + n);
  }
}

// one thread per image element; 72 CTAs of 256 threads
__global__ void __launch_bounds__(256) mlp_bwd_prep_kernel(const float* __restrict__ w, uint16_t* __restrict__ img) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= kImg16) return;
  constexpr int base[10] = {bW0, bW1, bW2, bW3, bT4, bT3, bT2, bT1, bT0, kImg16};
  constexpr int Ks[9] = {32, 64, 32, 64, 16, 64, 64, 16, 64};
  int mat = 0;
#pragma unroll
  for (int m = 1; m < 9; ++m)
    if (i >= base[m]) mat = m;
  const int K = Ks[mat], r = i - base[mat];
  const int n = r / K, k = r % K;
  uint32_t hi, lo;
  split_bf16x2(image_value(w, mat, n, k), 0.f, hi, lo);
  const int pos = base[mat] + canon16(n, k, K);
  img[pos] = (uint16_t)(hi & 0xFFFFu);
  img[kImg16 + pos] = (uint16_t)(lo & 0xFFFFu);
}

// ---- TMEM / shared-memory plan -------------------------------------------------------------------------------
constexpr uint32_t fColD = 0, fColAhi = 64, fColAlo = 96, fCtxCols = 128;
constexpr uint32_t fAccBase = 256, fAccCols = 80;
constexpr uint32_t fStashBase = 416, fStashCols = 32;  // per context: the hi words of h1, parked until dW1 needs them
constexpr uint32_t kLane16 = 16u << 16;  // TMEM address of the second M = 64 accumulator of a column range
constexpr uint32_t aW3 = 0, aW0 = 0 + kLane16, aW2 = 32 + kLane16, aW1 = 64, aW4 = 64 + kLane16;
constexpr uint32_t kLineBuf = kTile * 128;            // one [point][64 x bf16] line buffer: 16 KB
constexpr uint32_t kZBuf = kTile * 16;                // [point][8 x bf16]: 2 KB
constexpr uint32_t kCtxBytes = 4 * kLineBuf + 2 * kZBuf;  // X_hi | X_lo | Y_hi | Y_lo | Z_hi | Z_lo = 69,632
constexpr size_t kFusedSmemBytes = (size_t)kImgBytes + 2 * kCtxBytes + 1024;  // + slack for the 1024-byte alignment

// One layer: D[128 x N] = A[128 x K] . W[N x K]^T as 3 x (K/16) bf16 MMAs, small terms first.
template <int N, int K>
__device__ __forceinline__ void issue_layer16(uint32_t tmem, uint32_t w_hi_saddr, uint32_t w_lo_saddr) {
  constexpr uint32_t idesc = make_idesc_bf16(kTile, N);
  constexpr uint64_t kHiBits = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(((K / 8) * 128) >> 4) << 32) | (1ull << 46);
  const uint64_t d_hi = kHiBits | (uint64_t)((w_hi_saddr >> 4) & 0x3FFF);
  const uint64_t d_lo = kHiBits | (uint64_t)((w_lo_saddr >> 4) & 0x3FFF);
#pragma unroll
  for (int s = 0; s < K / 16; ++s) umma_ts_bf16(tmem + fColD, tmem + fColAlo + 8 * s, d_hi + 16 * s, idesc, s ? 1u : 0u);
#pragma unroll
  for (int s = 0; s < K / 16; ++s) umma_ts_bf16(tmem + fColD, tmem + fColAhi + 8 * s, d_lo + 16 * s, idesc, 1u);
#pragma unroll
  for (int s = 0; s < K / 16; ++s) umma_ts_bf16(tmem + fColD, tmem + fColAhi + 8 * s, d_hi + 16 * s, idesc, 1u);
}

// One weight gradient of one tile: acc[64 x N] += sum over the tile's 128 points of X[p][m] . Y[p][n].
// X: 128-byte swizzled lines; Y: the same, or (DENSE8) dense 16-byte rows.  8 K steps of 16 points.
template <int N, bool DENSE8>
__device__ __forceinline__ void issue_dw(uint32_t acc, uint32_t x_hi, uint32_t x_lo, uint32_t y_hi, uint32_t y_lo) {
  constexpr uint32_t idesc = make_idesc_bf16(64, N, 1, 1);
  const uint64_t xh = make_sdesc_mn_sw128(x_hi), xl = make_sdesc_mn_sw128(x_lo);
  const uint64_t yh = DENSE8 ? make_sdesc_mn_n8(y_hi) : make_sdesc_mn_sw128(y_hi);
  const uint64_t yl = DENSE8 ? make_sdesc_mn_n8(y_lo) : make_sdesc_mn_sw128(y_lo);
  constexpr uint64_t xs = 2048 >> 4, ys = (DENSE8 ? 256 : 2048) >> 4;
#pragma unroll
  for (int s = 0; s < 8; ++s) umma_ss_bf16(acc, xl + s * xs, yh + s * ys, idesc, 1u);
#pragma unroll
  for (int s = 0; s < 8; ++s) umma_ss_bf16(acc, xh + s * xs, yl + s * ys, idesc, 1u);
#pragma unroll
  for (int s = 0; s < 8; ++s) umma_ss_bf16(acc, xh + s * xs, yh + s * ys, idesc, 1u);
}

// 16 fp32 values -> 8 hi words + 8 lo words (value 2i in the low half of word i)
__device__ __forceinline__ void pack16(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16x2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
}

// 8 words = 16 features = two 16-byte chunks (chunk index c0, c0 + 1) of this thread's line; sw = point & 7
__device__ __forceinline__ void put_chunks(uint8_t* line_hi, uint8_t* line_lo, int sw, int c0, const uint32_t (&hi)[8],
                                           const uint32_t (&lo)[8]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int off = ((c0 + j) ^ sw) << 4;
    *reinterpret_cast<uint4*>(line_hi + off) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
    *reinterpret_cast<uint4*>(line_lo + off) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
  }
}

// what a 64-wide epilogue does with its packed words
enum : int { kToA = 1, kToX = 2, kToY = 4 };

struct Ctx {
  uint32_t tmem, row;          // this context's TMEM base; + this thread's lane
  uint32_t stash;              // this thread's lane of the context's stash columns
  uint8_t *xh, *xl, *yh, *yl;  // this thread's lines in the X / Y buffers
  int sw;
  uint64_t* bar_dw;            // completes when the weight-gradient MMAs issued last have read X / Y / Z
  uint32_t phase_dw;
  bool dw_pending;
};

// X, Y and Z may be overwritten only after the dW MMAs that read them are done.  They are issued BEHIND the layer
// MMAs of the same step and committed to their own barrier, so they execute while this context's threads already
// fetch the layer's accumulator from TMEM.
__device__ __forceinline__ void wait_dw(Ctx& c) {
  if (c.dw_pending) {
    mbar_wait(c.bar_dw, c.phase_dw);
    c.phase_dw ^= 1u;
    c.dw_pending = false;
  }
}

// reads this thread's 64 accumulator columns, applies ReLU (RELU: returns the positive mask; with `have_gate` the
// forward pass's mask decides instead of the recomputed sign) or the gate mask, packs to bf16 pairs and writes the
// words to the next A operand and / or this thread's X / Y line.  KEEP: also returns the words (the h1 stash).
template <bool RELU, bool GATES, int DEST, bool KEEP>
__device__ __forceinline__ uint64_t epilogue64(Ctx& c, uint64_t gate, uint32_t (&keep_hi)[32], uint32_t (&keep_lo)[32]) {
  constexpr bool have_gate = GATES;
  uint64_t mask = 0;
  float v[4][16];
  tmem_ld64(c.row + fColD, v);
  if (DEST & (kToX | kToY)) wait_dw(c);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (RELU && !have_gate) {
        if (v[q][i] > 0.f) mask |= (1ull << (16 * q + i));
        v[q][i] = fmaxf(v[q][i], 0.f);
      } else {
        v[q][i] = ((gate >> (16 * q + i)) & 1ull) ? v[q][i] : 0.f;
      }
    }
    uint32_t hi[8], lo[8];
    pack16(v[q], hi, lo);
    if (DEST & kToA) {
      tmem_st8(c.row + fColAhi + 8 * q, hi);
      tmem_st8(c.row + fColAlo + 8 * q, lo);
    }
    if (DEST & kToX) put_chunks(c.xh, c.xl, c.sw, 2 * q, hi, lo);
    if (DEST & kToY) put_chunks(c.yh, c.yl, c.sw, 2 * q, hi, lo);
    if (KEEP) {  // the h1 stash: hi words parked in spare TMEM columns, lo words kept in registers
      tmem_st8(c.stash + 8 * q, hi);
#pragma unroll
      for (int i = 0; i < 8; ++i) keep_lo[8 * q + i] = lo[i];
    }
  }
  return (RELU && !have_gate) ? mask : gate;
}

__device__ __forceinline__ void load_enc_row(float (&e)[32], int64_t p, int64_t N, const float* __restrict__ enc,
                                             int64_t enc_stride, int aligned) {
  if (p < N) {
    const float* erow = enc + p * enc_stride;
    if (aligned) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(erow) + k);
        e[4 * k] = f.x;
        e[4 * k + 1] = f.y;
        e[4 * k + 2] = f.z;
        e[4 * k + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) e[i] = __ldg(erow + i);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) e[i] = 0.f;
  }
}

// one step boundary: this thread's tcgen05.st and shared-memory stores are made visible, the context's 128
// threads meet, its leader issues `issue()` + one commit, everybody waits for the commit.
template <class F>
__device__ __forceinline__ void step(int sync_id, bool leader_warp, uint64_t* bar, uint32_t& phase, F&& issue) {
  wait_st();
  fence_async_smem();
  fence_before_sync();
  ctx_sync(sync_id);
  if (leader_warp) {  // warp-uniform branch; one elected lane issues
    if (elect_one()) {
      fence_after_sync();
      issue();
    }
    __syncwarp();
  }
  mbar_wait(bar, phase);
  phase ^= 1u;
  fence_after_sync();
}

template <bool HAVE_GATES>
__global__ void __launch_bounds__(2 * kTile, 1)
mlp_tc_bwd_fused_kernel(const float* __restrict__ enc, int64_t enc_stride, const float* __restrict__ views,
                        int64_t views_stride, int64_t pts_per_view, const uint16_t* __restrict__ images,
                        const uint8_t* __restrict__ keep, const uint32_t* __restrict__ gates,
                        const float* __restrict__ dout, int64_t N, float* __restrict__ d_enc,
                        float* __restrict__ partials, int aligned, int ablate) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bars[5];  // [0], [1]: layer MMAs of context 0 / 1; [2], [3]: their dW MMAs; [4]: images
  __shared__ uint32_t tmem_slot;
  // warp-uniform ids, broadcast from lane 0 so that the compiler keeps everything derived from them (TMEM and
  // shared-memory operand addresses, descriptors) in uniform registers: the MMA issue sequence is then a handful
  // of uniform-datapath instructions per tcgen05.mma instead of a per-instruction vector -> uniform "waterfall".
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int ctx = wid >> 2, warp = wid & 3, lane = threadIdx.x & 31, t = warp * 32 + lane;
  const int sync_id = 1 + ctx;
  const bool leader_warp = (warp == 0);
  uint64_t* bar = &bars[ctx];

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    // the nine weight images (hi + lo), prepared once per launch by mlp_bwd_prep_kernel: one TMA bulk copy
    mbar_arrive_expect_tx(&bars[4], kImgBytes);
    bulk_g2s(smem, images, kImgBytes, &bars[4]);
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // the CTA owns all 512 columns, so its allocation starts at TMEM address 0: addresses are compile-time offsets
  if (tmem_slot != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;
  Ctx c;
  c.tmem = tmem_base + (uint32_t)ctx * fCtxCols;
  c.row = c.tmem + ((uint32_t)(warp * 32) << 16);
  c.stash = tmem_base + fStashBase + (uint32_t)ctx * fStashCols + ((uint32_t)(warp * 32) << 16);
  uint8_t* cbase = smem + kImgBytes + (uint32_t)ctx * kCtxBytes;
  c.xh = cbase + t * 128;
  c.xl = cbase + kLineBuf + t * 128;
  c.yh = cbase + 2 * kLineBuf + t * 128;
  c.yl = cbase + 3 * kLineBuf + t * 128;
  c.sw = t & 7;
  c.bar_dw = &bars[2 + ctx];
  c.phase_dw = 0;
  c.dw_pending = false;
  uint8_t* zh = cbase + 4 * kLineBuf + t * 16;
  uint8_t* zl = zh + kZBuf;
  const uint32_t s_hi = smem_u32(smem), s_lo = s_hi + kImg16 * 2;
  const uint32_t sX_hi = smem_u32(cbase), sX_lo = sX_hi + kLineBuf, sY_hi = sX_lo + kLineBuf, sY_lo = sY_hi + kLineBuf;
  const uint32_t sZ_hi = sY_lo + kLineBuf, sZ_lo = sZ_hi + kZBuf;
  const uint32_t acc = tmem_base + fAccBase + (uint32_t)ctx * fAccCols;        // MMA view (lane 0)
  const uint32_t acc_row = acc + ((uint32_t)(warp * 32) << 16);                // this thread's lane
  uint32_t phase = 0;

  {  // accumulators start at zero: every dW MMA accumulates
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
    for (int c0 = 0; c0 < (int)fAccCols; c0 += 16) tmem_st16(acc_row + c0, z);
  }
  mbar_wait(&bars[4], 0);  // weight images in shared memory

  const int64_t n_tiles = (N + kTile - 1) / kTile;
  const int64_t tile_step = (int64_t)gridDim.x * 2;
  const int64_t tile0 = (int64_t)blockIdx.x * 2 + ctx;
  TileInputs cur;
  float4 go_cur = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tile0 < n_tiles) {
    const int64_t p0 = tile0 * kTile + t;
    load_tile_inputs(cur, p0, N, enc, enc_stride, views, views_stride, pts_per_view, keep, aligned);
    if (p0 < N) go_cur = __ldg(reinterpret_cast<const float4*>(dout) + p0);
  }

  for (int64_t tile = tile0; tile < n_tiles; tile += tile_step) {
    const int64_t p = tile * kTile + t;
    const bool valid = p < N;
    // ---- S0: A = hash features (K = 32)
    {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        pack16(cur.e[h], hi, lo);
        tmem_st8(c.row + fColAhi + 8 * h, hi);
        tmem_st8(c.row + fColAlo + 8 * h, lo);
      }
    }
    float vsh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) vsh[i] = cur.v[i];
    const float4 go = go_cur;
    const float dsigma = (valid && cur.keep != 0) ? go.w : 0.f;
    // the forward pass's ReLU gates, when the caller kept them
    uint64_t g1 = 0, g3 = 0, g4 = 0;
    if (HAVE_GATES && valid) {
      const uint2* gp = reinterpret_cast<const uint2*>(gates + p * 6);
      const uint2 a = __ldg(gp), b = __ldg(gp + 1), d = __ldg(gp + 2);
      g1 = ((uint64_t)a.y << 32) | a.x;
      g3 = ((uint64_t)b.y << 32) | b.x;
      g4 = ((uint64_t)d.y << 32) | d.x;
    }
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<64, 32>(c.tmem, s_hi + bW0 * 2, s_lo + bW0 * 2);
      umma_commit(bar);
    });
    // ---- E0: h1 = relu(.) -> A (K = 64); the packed words are kept for dW1
    uint32_t h1hi[32], h1lo[32];  // h1hi: unused (the hi words live in TMEM)
    const uint64_t m1 = epilogue64<true, HAVE_GATES, kToA, true>(c, g1, h1hi, h1lo);
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<16, 64>(c.tmem, s_hi + bW1 * 2, s_lo + bW1 * 2);
      umma_commit(bar);
    });
    // ---- E1: h2 = [sigma | geo]; c = [sh(16) | geo(15) | 0] -> A (K = 32); the geo words are kept for dW2
    uint32_t ghi[8], glo[8];
    {
      float h2[16];
      tmem_ld16(c.row + fColD, h2);
      float g[16];
#pragma unroll
      for (int i = 0; i < 15; ++i) g[i] = h2[1 + i];
      g[15] = 0.f;
      uint32_t hi[8], lo[8];
      pack16(vsh, hi, lo);
      tmem_st8(c.row + fColAhi, hi);
      tmem_st8(c.row + fColAlo, lo);
      pack16(g, ghi, glo);
      tmem_st8(c.row + fColAhi + 8, ghi);
      tmem_st8(c.row + fColAlo + 8, glo);
    }
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<64, 32>(c.tmem, s_hi + bW2 * 2, s_lo + bW2 * 2);
      umma_commit(bar);
    });
    // ---- E2: h3 -> A (K = 64) and -> Y (N side of dW3)
    uint32_t dummy_hi[32], dummy_lo[32];
    const uint64_t m3 = epilogue64<true, HAVE_GATES, kToA | kToY, false>(c, g3, dummy_hi, dummy_lo);
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<64, 64>(c.tmem, s_hi + bW3 * 2, s_lo + bW3 * 2);
      umma_commit(bar);
    });
    // ---- E3: h4 -> X (M side of dW4); drgb -> Z and -> A (K = 16, columns 3..15 zero)
    const uint64_t m4 = epilogue64<true, HAVE_GATES, kToX, false>(c, g4, dummy_hi, dummy_lo);
    {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hi[i] = lo[i] = 0u;
      split_bf16x2(go.x, go.y, hi[0], lo[0]);
      split_bf16x2(go.z, 0.f, hi[1], lo[1]);
      tmem_st8(c.row + fColAhi, hi);
      tmem_st8(c.row + fColAlo, lo);
      wait_dw(c);
      *reinterpret_cast<uint4*>(zh) = make_uint4(hi[0], hi[1], 0u, 0u);
      *reinterpret_cast<uint4*>(zl) = make_uint4(lo[0], lo[1], 0u, 0u);
    }
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<64, 16>(c.tmem, s_hi + bT4 * 2, s_lo + bT4 * 2);            // dh4 = drgb . W4
      umma_commit(bar);
      issue_dw<8, true>(acc + aW4, sX_hi, sX_lo, sZ_hi, sZ_lo);                  // dW4^T += h4^T . drgb
      umma_commit(c.bar_dw);
    });
    c.dw_pending = true;
    // ---- E4: dz4 = dh4 . [h4 > 0] -> A (K = 64) and -> X
    epilogue64<false, false, kToA | kToX, false>(c, m4, dummy_hi, dummy_lo);
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<64, 64>(c.tmem, s_hi + bT3 * 2, s_lo + bT3 * 2);            // dh3 = dz4 . W3
      umma_commit(bar);
      issue_dw<64, false>(acc + aW3, sX_hi, sX_lo, sY_hi, sY_lo);                // dW3 += dz4^T . h3
      umma_commit(c.bar_dw);
    });
    c.dw_pending = true;
    // ---- E5: dz3 -> A (K = 64) and -> X; c = [sh | geo] -> Y (32 features)
    epilogue64<false, false, kToA | kToX, false>(c, m3, dummy_hi, dummy_lo);
    {
      uint32_t hi[8], lo[8];
      pack16(vsh, hi, lo);
      put_chunks(c.yh, c.yl, c.sw, 0, hi, lo);
      put_chunks(c.yh, c.yl, c.sw, 2, ghi, glo);
    }
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<16, 64>(c.tmem, s_hi + bT2 * 2, s_lo + bT2 * 2);            // dgeo = (dz3 . W2)[16:31]
      umma_commit(bar);
      issue_dw<32, false>(acc + aW2, sX_hi, sX_lo, sY_hi, sY_lo);                // dW2 += dz3^T . c
      umma_commit(c.bar_dw);
    });
    c.dw_pending = true;
    // ---- E6: dh2 = [dsigma | dgeo] -> A (K = 16) and -> Y (16 features); h1 -> X
    {
      float dg[16], v[16];
      tmem_ld16(c.row + fColD, dg);
      v[0] = dsigma;
#pragma unroll
      for (int i = 0; i < 15; ++i) v[1 + i] = dg[i];
      uint32_t hi[8], lo[8];
      pack16(v, hi, lo);
      tmem_st8(c.row + fColAhi, hi);
      tmem_st8(c.row + fColAlo, lo);
      wait_dw(c);
      put_chunks(c.yh, c.yl, c.sw, 0, hi, lo);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tmem_ld8w(c.stash + 8 * q, hi);
#pragma unroll
        for (int i = 0; i < 8; ++i) lo[i] = h1lo[8 * q + i];
        put_chunks(c.xh, c.xl, c.sw, 2 * q, hi, lo);
      }
    }
    // the hash features again (N side of dW0, needed one step from now)
    float e2[32];
    load_enc_row(e2, p, N, enc, enc_stride, aligned);
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<64, 16>(c.tmem, s_hi + bT1 * 2, s_lo + bT1 * 2);            // dh1 = dh2 . W1
      umma_commit(bar);
      issue_dw<16, false>(acc + aW1, sX_hi, sX_lo, sY_hi, sY_lo);                // dW1^T += h1^T . dh2
      umma_commit(c.bar_dw);
    });
    c.dw_pending = true;
    // ---- E7: dz1 -> A (K = 64) and -> X; hash features -> Y (32 features)
    epilogue64<false, false, kToA | kToX, false>(c, m1, dummy_hi, dummy_lo);
    {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = e2[16 * h + i];
        pack16(v, hi, lo);
        put_chunks(c.yh, c.yl, c.sw, 2 * h, hi, lo);
      }
    }
    // the next tile's inputs: in flight while T0 runs and d_enc is written
    go_cur = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tile + tile_step < n_tiles) {
      const int64_t pn = (tile + tile_step) * kTile + t;
      load_tile_inputs(cur, pn, N, enc, enc_stride, views, views_stride, pts_per_view, keep, aligned);
      if (pn < N) go_cur = __ldg(reinterpret_cast<const float4*>(dout) + pn);
    }
    step(sync_id, leader_warp, bar, phase, [&] {
      issue_layer16<32, 64>(c.tmem, s_hi + bT0 * 2, s_lo + bT0 * 2);            // d_enc = dz1 . W0
      umma_commit(bar);
      issue_dw<32, false>(acc + aW0, sX_hi, sX_lo, sY_hi, sY_lo);                // dW0 += dz1^T . in
      umma_commit(c.bar_dw);
    });
    c.dw_pending = true;
    // ---- E8: d_enc out
#pragma unroll
    for (int c0 = 0; c0 < kIn; c0 += 16) {
      float v[16];
      tmem_ld16(c.row + fColD + c0, v);
      if (valid) {
        float* drow = d_enc + p * kIn + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(drow)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }

  // ---- drain the last dW0 group, then flush this context's accumulators
  wait_st();
  fence_before_sync();
  ctx_sync(sync_id);
  wait_dw(c);
  fence_after_sync();
  // ---- the CTA's weight gradient -> partials[blockIdx.x][9344]; mlp_dw_reduce_kernel sums the rows.  (Atomics on
  // dweights from every context cost 37-42 us when the 296 contexts of a small batch finish together: 2.7 M atomics on
  // 9344 addresses.)  Context 1 parks its 80 values per thread in shared memory -- the operand buffers are free now,
  // and thread t of either context holds the same (row, column) elements -- context 0 adds its own and stores.
  float* stage = reinterpret_cast<float*>(smem + kImgBytes) + t;   // [80][128]: conflict-free
  __syncthreads();  // both contexts have drained their MMAs: context 0's operand buffers may be overwritten
  if (ctx == 1 && !(ablate & 1)) {
    float v[16];
#pragma unroll
    for (int c0 = 0; c0 < (int)fAccCols; c0 += 16) {
      tmem_ld16(acc_row + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) stage[(c0 + i) * kTile] = v[i];
    }
  }
  __syncthreads();
  if (ctx == 0 && !(ablate & 1)) {
    // M = 64 accumulator rows: row m sits in lane (m % 16) + 32 (m / 16) of its lane half; a thread with
    // lane < 16 holds row 16 warp + lane of the first accumulator of each column range, the others row
    // 16 warp + lane - 16 of the second one.
    float* part = partials + (size_t)blockIdx.x * kGTotal;
    const bool second = lane >= 16;
    const int m = warp * 16 + (lane & 15);
    float v[16];
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      tmem_ld16(acc_row + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += stage[(c0 + i) * kTile];
      if (!second) {                                  // dW3[j = m][k = c0 ..]
        float4* dst = reinterpret_cast<float4*>(part + kG3 + m * 64 + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else if (c0 < 32) {                           // dW0[j = m][k = c0 ..]
        float4* dst = reinterpret_cast<float4*>(part + kG0 + m * 32 + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {                                        // dW2[j = m][c = c0 - 32 ..], rows of 31
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 - 32 + i < 31) part[kG2 + m * 31 + (c0 - 32 + i)] = v[i];
      }
    }
    tmem_ld16(acc_row + 64, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = v[i] + stage[(64 + i) * kTile];
      if (!second) part[kG1 + i * 64 + m] = x;        // dW1^T[k = m][j = i]
      else if (i < 3) part[kG4 + i * 64 + m] = x;     // dW4^T[k = m][c = i]
    }
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

// dweights[j] += sum over the CTAs' partial rows; 32 float4 outputs per CTA, its 8 warps split the rows
__global__ void __launch_bounds__(256)
mlp_dw_reduce_kernel(const float* __restrict__ partials, int n_rows, float* __restrict__ dweights) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j4 = blockIdx.x * 32 + lane;  // float4 index into the 9344 floats
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j4 < kGTotal / 4) {
#pragma unroll 4
    for (int r = warp; r < n_rows; r += 8) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partials + (size_t)r * kGTotal) + j4);
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && j4 < kGTotal / 4) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      s.x += red[w][lane].x;
      s.y += red[w][lane].y;
      s.z += red[w][lane].z;
      s.w += red[w][lane].w;
    }
    float* d = dweights + 4 * j4;   // accumulate: the caller's buffer may already hold the other pass's gradient
    atomicAdd(d, s.x);
    atomicAdd(d + 1, s.y);
    atomicAdd(d + 2, s.z);
    atomicAdd(d + 3, s.w);
  }
}

}  // namespace tc

// weight images + one partial weight-gradient row per CTA
int64_t mlp_tc_bwd_fused_workspace_bytes() {
  return (int64_t)tc::kImgBytes + (int64_t)sm_count() * tc::kGTotalBytes;
}

int mlp_tc_bwd_fused(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride, int64_t pts_per_view,
                     const float* weights, const uint8_t* keep, const uint32_t* gates, const float* dout, int64_t N,
                     float* d_enc, float* dweights, float* workspace, int aligned, cudaStream_t stream) {
  static thread_local int done_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, "cudaGetDevice");
  if (done_dev != dev) {
    e = cudaFuncSetAttribute(tc::mlp_tc_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)tc::kFusedSmemBytes);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_tc_bwd_fused_kernel)");
    e = cudaFuncSetAttribute(tc::mlp_tc_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)tc::kFusedSmemBytes);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mlp_tc_bwd_fused_kernel)");
    done_dev = dev;
  }
  uint16_t* images = reinterpret_cast<uint16_t*>(workspace);
  int rc = 0;
  if (!(g_mlp_dw_ablate & 2)) {
    tc::mlp_bwd_prep_kernel<<<(tc::kImg16 + 255) / 256, 256, 0, stream>>>(weights, images);
    rc = check_launch("mlp_bwd_prep_kernel");
    if (rc) return rc;
  }
  const int64_t tiles = (N + tc::kTile - 1) / tc::kTile;
  const int64_t cap = (int64_t)sm_count();
  const int64_t want = (tiles + 1) / 2;  // two tile contexts per CTA
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + tc::kImgBytes);
  if (gates != nullptr)
    tc::mlp_tc_bwd_fused_kernel<true><<<grid, 2 * tc::kTile, tc::kFusedSmemBytes, stream>>>(
        enc, enc_stride, views, views_stride, pts_per_view, images, keep, gates, dout, N, d_enc, partials, aligned,
        g_mlp_dw_ablate);
  else
    tc::mlp_tc_bwd_fused_kernel<false><<<grid, 2 * tc::kTile, tc::kFusedSmemBytes, stream>>>(
        enc, enc_stride, views, views_stride, pts_per_view, images, keep, gates, dout, N, d_enc, partials, aligned,
        g_mlp_dw_ablate);
  if ((rc = check_launch("mlp_tc_bwd_fused_kernel"))) return rc;
  if (g_mlp_dw_ablate & 1) return 0;
  tc::mlp_dw_reduce_kernel<<<(tc::kGTotal / 4 + 31) / 32, 256, 0, stream>>>(partials, (int)grid, dweights);
  return check_launch("mlp_dw_reduce_kernel");
}

}  // namespace hn
