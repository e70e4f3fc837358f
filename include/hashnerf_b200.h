/*
 * hashnerf_b200.h -- C ABI of libhashnerf_b200.so, the sm_100a implementation of the HashNeRF hot path
 *                    HashEmbedder -> SHEncoder -> NeRFSmall -> raw2outputs / sample_pdf.
 *
 * The reference (mache102/HashNeRF-pytorch) is pure Python and has no FFI layer; the "interface each
 * entry point replaces" is therefore the Python function whose ATen op chain it fuses.  Citations are
 * file:line into the reference checkout.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to densely packed row-major data unless a stride is passed;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it: no allocation,
 *     no synchronisation, no pointer retained after return (scratch is passed in by the caller);
 *   - return value: 0 on success, a positive cudaError_t, or HN_EINVAL for a rejected argument;
 *     hn_last_error_string() describes the last non-zero return of the calling thread;
 *   - nothing throws, exits or prints.
 */
#ifndef HASHNERF_B200_H_
#define HASHNERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HN_API __attribute__((visibility("default")))
#else
#define HN_API
#endif

#define HN_EINVAL (-22)
#define HN_MAX_LEVELS 32
#define HN_ABI_VERSION 2

/* ---- library ---------------------------------------------------------------------------------- */
HN_API int hn_abi_version(void);
HN_API const char* hn_last_error_string(void);
/* Multiprocessor count / compute capability of the current device (used to size persistent grids). */
HN_API int hn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Launch-shape knobs for profiling sweeps and A/B runs (not part of the numerical contract; results stay within
 * the stated tolerances for every setting).  Keys:
 *   "hash_fwd_lpg", "hash_bwd_lpg"   levels per thread: 1,2,4,8,16; 0 = heuristic
 *   "hash_bwd_agg"                   warp-aggregated scatter: -1 = sorted / ordered points only, 0 = never, 1 = always
 *   "hash_agg_max_heads"             aggregate a level only if a warp's 32 lanes form at most this many runs (24)
 *   "hash_level_major"               -1 = level-major grid for caller-ordered points, tile-major for sorted; 0/1 force
 *   "hash_sort_two_level"            1 = two-level counting sort, one global cursor per bin (default), 0 = single-pass sort
 *   "hash_div_hoist"                 1 = hoisted-reciprocal cell-index division (default), 0 = per-point true division
 *   "mlp_impl"                       1 = tcgen05 tensor-core MLP (default), 0 = FFMA fp32 MLP
 *   "mlp_bwd_impl"                   tcgen05 backward: 1 = one fused kernel, bf16 hi+lo operands, weight gradients
 *                                    formed on chip (default); 0 = two 3xTF32 kernels around an HBM workspace.
 *                                    hn_mlp_bwd_workspace_bytes() follows the selection: query it after changing it
 *   "mlp_dw_nbuf"                    weight-gradient kernel: 1 = two CTAs/SM, one staging buffer (default); 2 = one
 *                                    CTA/SM, two buffers
 *   "dp_grid_per_sm"                 CTAs per SM of hn_dp_reduce_update (default 8)
 *   "mlp_fwd_one_cta", "mlp_dw_ablate"   profiling only (occupancy / phase ablations; the latter breaks results)
 * Unknown keys return HN_EINVAL.  The knobs are PROCESS-GLOBAL and unsynchronised: set them before the threads that
 * launch kernels start (they exist for sweeps and A/B runs, not for per-call configuration). */
HN_API int hn_set_tuning(const char* key, int value);

/* ---- (a3) spatial hash : embedding/hash_encoding.py:112-128 ------------------------------------ */
/* hashed[i] = (XOR_d coords[i*dim+d] * prime_d) & (2^log2T - 1); dim <= 7, int64 in / int64 out as
 * the reference; used by loss.py:29 on arbitrary integer coordinates. */
HN_API int hn_spatial_hash(const int64_t* coords, int64_t n, int dim, int log2T, int64_t* hashed, void* stream);

/* ---- (a2) voxel vertices of every level : embedding/hash_encoding.py:59-82 --------------------- */
/* Debug/parity entry point exposing the intermediates the fused encoder never materialises:
 * hashed [L,N,8] int64 (corner c = 4i+2j+k), vmin/vmax [L,N,3] f32 (any may be NULL).
 * bbox = {min.x,min.y,min.z,max.x,max.y,max.z}; resolutions = floor(base*b^l) as f32 [L]. */
HN_API int hn_voxel_vertices(const float* x, const float* bbox, const float* resolutions, int64_t N, int L,
                      int log2T, int64_t* hashed, float* vmin, float* vmax, void* stream);

/* ---- (a4-a6) multiresolution hash encoding : embedding/hash_encoding.py:84-110, 130-163 -------- */
/* tables: [L, 2^log2T, F] f32 (level l's nn.Embedding.weight is the l-th slab).  F in {1, 2, 4}.
 * out: [N, L*F] f32, level-major features.  keep: [N] uint8, the mask forward() returns
 * (last level's in-box test on the already-clamped coordinates; may be NULL).
 * tables, out (and dy, dtables, xs4 of the calls below) must be 16-byte aligned: the kernels use vector accesses;
 * a violation returns HN_EINVAL instead of faulting.  N * ceil(L / levels-per-thread) must fit a 1-D grid (2^31-1). */
HN_API int hn_hash_encode_fwd(const float* x, const float* tables, const float* bbox, const float* resolutions,
                       int64_t N, int L, int F, int log2T, float* out, uint8_t* keep, void* stream);
/* Autograd of the above w.r.t. the tables: dtables[l, h_c, :] += dy[p, l*F:(l+1)*F] * W_c(p) for the 8
 * corners.  ACCUMULATES into dtables (caller zeroes).  No gradient w.r.t. x exists on this path. */
HN_API int hn_hash_encode_bwd(const float* x, const float* dy, const float* bbox, const float* resolutions,
                       int64_t N, int L, int F, int log2T, float* dtables, void* stream);

/* Same as hn_hash_encode_bwd for points the caller knows to be spatially coherent in their given order
 * (consecutive samples of a ray): lanes of a warp that fall in the same voxel are summed before the atomics. */
HN_API int hn_hash_encode_bwd_ordered(const float* x, const float* dy, const float* bbox, const float* resolutions,
                                      int64_t N, int L, int F, int log2T, float* dtables, void* stream);

/* Coherent ("sorted") variant of the two calls above.  hn_hash_sort_points bins the points into a
 * grid_res^3 grid over the bbox (counting sort, x fastest) and writes xs4[i] = (x, y, z, bit-cast original
 * row) for the i-th point in cell order; the *_sorted calls process that order -- gathers coalesce / hit L1,
 * the scatter sums lanes that share a voxel in-warp before issuing atomics -- and read dy / write out, keep at
 * the ORIGINAL rows, so results are indistinguishable from the plain calls (forward bit-identical, backward
 * up to atomic summation order).  workspace: hn_hash_sort_workspace_bytes(N, grid_res) bytes, 16-byte aligned;
 * xs4: [N] float4.  N < 2^32.  The order of the points INSIDE one grid cell (and the workspace contents) may differ
 * from call to call: slots are taken with atomics. */
HN_API int64_t hn_hash_sort_workspace_bytes(int64_t N, int grid_res);
HN_API int hn_hash_sort_points(const float* x, const float* bbox, int64_t N, int grid_res, void* workspace,
                               float* xs4, void* stream);
HN_API int hn_hash_encode_fwd_sorted(const float* xs4, const float* tables, const float* bbox,
                                     const float* resolutions, int64_t N, int L, int F, int log2T, float* out,
                                     uint8_t* keep, void* stream);
HN_API int hn_hash_encode_bwd_sorted(const float* xs4, const float* dy, const float* bbox,
                                     const float* resolutions, int64_t N, int L, int F, int log2T, float* dtables,
                                     void* stream);
/* The scatter of hn_hash_encode_bwd_sorted restricted to levels [level_begin, level_end): dy and dtables are the
 * FULL [N, L*F] / [L, 2^log2T, F] arrays.  Calling it over a partition of [0, L) equals one full call; it exists so
 * that a data-parallel caller can all-reduce the table-gradient slab of one level bucket while the next bucket is
 * still being scattered (hn_b200/dp.py: BucketedTableReducer).  Ranges aligned to 4 levels keep the
 * kernel's 4-levels-per-thread mapping; other ranges run with fewer levels per thread. */
HN_API int hn_hash_encode_bwd_sorted_levels(const float* xs4, const float* dy, const float* bbox,
                                            const float* resolutions, int64_t N, int L, int F, int log2T,
                                            float* dtables, int level_begin, int level_end, void* stream);

/* ---- (a7) spherical harmonics : embedding/spherical_harmonic.py:65-103 ------------------------- */
/* dirs [N,3] -> out [N, degree^2], 1 <= degree <= 5. */
HN_API int hn_sh_encode(const float* dirs, int64_t N, int degree, float* out, void* stream);

/* ---- (a8) NeRFSmall : models.py:151-174 as instantiated at run_nerf_helpers.py:79-84 ----------- */
/* Geometry is fixed to the instantiated network: input_ch 32, input_ch_views 16, hidden 64,
 * geo_feat_dim 15, 2 sigma layers, 3 colour layers, no bias.
 * weights: the five nn.Linear.weight matrices ([out,in] row-major) packed back to back:
 *   W0[64,32] W1[16,64] W2[64,31] W3[64,64] W4[3,64]  = 9344 floats  (HN_MLP_PARAMS).
 * enc:   row p at enc + p*enc_stride (32 floats);   views: row (p / pts_per_view) at
 * views + (p / pts_per_view)*views_stride (16 floats).  pts_per_view = 1 reproduces
 * NeRFSmall.forward(x[N,48]) on a split view of x; pts_per_view = S evaluates SH once per ray
 * (fuses run_nerf_helpers.py:219-222).  out: [N,4] = (rgb_raw[3], sigma).
 * keep (may be NULL): sigma is written as 0 where keep[p] == 0 (run_nerf_helpers.py:225).
 * gates (may be NULL; [N][6] uint32, 8-byte aligned): receives, per point, the bit masks of strictly positive
 *   pre-activations of the three ReLU layers (h1, h3, h4; 64 bits each, low word first) -- the part of the forward
 *   state autograd would keep (models.py:157,169 F.relu).  Pass them to hn_mlp_bwd so that its ReLU derivative is
 *   the forward's even where the backward recomputes activations in another precision. */
#define HN_MLP_PARAMS 9344
#define HN_MLP_GATE_WORDS 6
HN_API int hn_mlp_fwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride,
               int64_t pts_per_view, const float* weights, const uint8_t* keep, int64_t N, float* out,
               uint32_t* gates, void* stream);
/* Backward: recomputes activations from (enc, views, weights).  dout [N,4].  d_enc [N,32] (dense, written),
 * dweights [9344] (ACCUMULATED).  No gradient w.r.t. views on this path (directions are data).
 * gates (may be NULL): the masks hn_mlp_fwd wrote for the same inputs; NULL = take the sign of the recomputed
 * pre-activations (exact for the fp32 / 3xTF32 implementations, within 1e-5 of a kink for the fused bf16x2 one).
 * workspace: caller-provided scratch of hn_mlp_bwd_workspace_bytes(N) bytes, 16-byte aligned (fused implementation:
 * the bf16 weight images + one partial weight-gradient row per CTA, summed into dweights by a second small kernel --
 * no atomics from the main kernel, and a summation order that is the same on every call). */
HN_API int64_t hn_mlp_bwd_workspace_bytes(int64_t N);
HN_API int hn_mlp_bwd(const float* enc, int64_t enc_stride, const float* views, int64_t views_stride,
               int64_t pts_per_view, const float* weights, const uint8_t* keep, const uint32_t* gates,
               const float* dout, int64_t N, float* d_enc, float* dweights, float* workspace, void* stream);

/* ---- (a10) raw2outputs : run_nerf_helpers.py:577-628 ------------------------------------------- */
/* raw [R,S,4], z [R,S], rays_d [R,3], noise [R,S] or NULL (already scaled by raw_noise_std).
 * Outputs: rgb [R,3], disp [R], acc [R], weights [R,S], depth [R], entropy [R]. */
HN_API int hn_composite_fwd(const float* raw, const float* z, const float* rays_d, const float* noise, int64_t R,
                     int S, int white_bkgd, float* rgb, float* disp, float* acc, float* weights, float* depth,
                     float* entropy, void* stream);
/* Any upstream gradient pointer may be NULL (= zero).  d_raw [R,S,4] is written. */
HN_API int hn_composite_bwd(const float* raw, const float* z, const float* rays_d, const float* noise, int64_t R,
                     int S, int white_bkgd, const float* d_rgb, const float* d_disp, const float* d_acc,
                     const float* d_weights, const float* d_depth, const float* d_entropy, float* d_raw,
                     void* stream);

/* ---- (a11) sample_pdf : run_nerf_helpers.py:264-307 -------------------------------------------- */
/* bins [R,nb], weights [R,nb-1], u [R,Ni] or NULL for det (u_i = linspace(0,1,Ni) passed as u_det [Ni]);
 * samples [R,Ni]. */
HN_API int hn_sample_pdf(const float* bins, const float* weights, const float* u, const float* u_det, int64_t R,
                  int nb, int Ni, float* samples, void* stream);
/* Fused form of the resampling block of render_rays (:547-552, :568): mids, sample_pdf on weights[:,1:-1],
 * sort(cat(z, samples)) and std(samples).  z, weights [R,S]; u [R,Ni] or NULL (then u_det [Ni]); outputs
 * samples [R,Ni], merged [R,S+Ni], z_std [R] (may be NULL).  S + Ni <= 2048. */
HN_API int hn_resample(const float* z, const float* weights, const float* u, const float* u_det, int64_t R, int S,
                       int Ni, float* samples, float* merged, float* z_std, void* stream);
/* out[r,:] = sort(cat(a[r,:na], b[r,:nb]))  (run_nerf_helpers.py:551); na+nb <= 2048. */
HN_API int hn_sort_concat_rows(const float* a, int na, const float* b, int nb, int64_t R, float* out, void* stream);

/* ---- (a12) ray-marching set-up : run_nerf_helpers.py:514-538 ----------------------------------- */
/* z[r,s] from near/far ([R] each, element stride nf_stride), t_vals [S] (= torch.linspace(0,1,S)),
 * optional stratified jitter t_rand [R,S] (NULL = none). */
HN_API int hn_coarse_z(const float* near, const float* far, int64_t nf_stride, const float* t_vals,
                const float* t_rand, int64_t R, int S, int lindisp, float* z, void* stream);
/* pts[r,s,:] = o[r,:] + d[r,:]*z[r,s]  (mul then add, separately rounded).  o/d rows have ray_stride. */
HN_API int hn_ray_points(const float* rays_o, const float* rays_d, int64_t ray_stride, const float* z, int64_t R,
                  int S, float* pts, void* stream);

/* ---- section 8f "next" row 1: RAdam : radam.py:34-92 ----------------------------------------------- */
/* One fused pass over a flat span of n parameters: moments (:58-59), weight decay (:82-83) and the
 * rectified-Adam or degenerated-SGD update (:84-90).  step_size and the branch (`mode`: 0 = moments
 * only, 1 = adaptive, 2 = SGD-like) are computed on the host from the step count as in :62-78.
 * The gradient is multiplied by grad_scale first (1/world_size after a summed all-reduce).  lr, weight_decay and
 * step_size are doubles: the reference multiplies them as python floats before ATen rounds the product (:83,:85).
 * zero_grad != 0 clears g in the same pass (optimizer.zero_grad(), run_nerf.py:612, folded in).  16-byte
 * vector accesses when p, g, m and v are all 16-byte aligned, scalar otherwise. */
HN_API int hn_radam_step(float* p, float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps,
                         double lr, double weight_decay, double step_size, int mode, float grad_scale,
                         int zero_grad, void* stream);

/* ---- section 8f "next" row 3: ray generation / packing ------------------------------------------------- */
/* ray_util.py:62-80: rays_d[H,W,3] for a pinhole camera (fx, fy, cx, cy) and a camera-to-world matrix c2w
 * (3 rows of >= 3 floats, row stride in floats).  rays_o is the broadcast translation column: no kernel. */
HN_API int hn_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w,
                       int64_t c2w_row_stride, float* rays_d, void* stream);
/* run_nerf_helpers.py:344-366: out[r] = (o, d, near, far[, viewdirs/|viewdirs|]); rows of o, d, viewdirs have
 * the given strides (in floats); viewdirs may be NULL (row width 8 instead of 11). */
HN_API int hn_pack_rays(const float* rays_o, int64_t o_stride, const float* rays_d, int64_t d_stride,
                        const float* viewdirs, int64_t vd_stride, float near, float far, int64_t R, float* out,
                        void* stream);

/* ray_util.py:96-142 (get_ndc_rays): shift the origins to the near plane and map (o, d) to normalised device
 * coordinates; rows of rays_o / rays_d have the given strides (floats); out_o / out_d are dense [R,3].  focal and
 * near are doubles: the reference folds them into python-float scalars before ATen rounds them. */
HN_API int hn_ndc_rays(int H, int W, double focal, double near, const float* rays_o, int64_t o_stride,
                       const float* rays_d, int64_t d_stride, int64_t R, float* out_o, float* out_d, void* stream);
/* Opt-in on-device training batcher replacing run_nerf.py:576-605 + run_nerf_helpers.py:344-366: n_rand DISTINCT
 * pixels (a keyed Feistel permutation, i.e. sampling without replacement like np.random.choice at :600) of a window
 * of image `step[0]`; writes rays [n_rand,11] = (o, d, near, far, d/|d|) with the ray arithmetic of hn_get_rays /
 * hn_pack_rays, target [n_rand,3] = images[img,row,col,0:3], and (pix != NULL) the chosen (row, col) pairs.
 * images [n_img,H,W,channels] and poses [n_img,3,4] stay resident on the device; step = int32[6] ON THE DEVICE
 * {image index, seed, row0, col0, win_h, win_w} so that the launch can be captured in a CUDA graph. */
HN_API int hn_sample_rays(const float* images, int channels, const float* poses, int H, int W, float fx, float fy,
                          float cx, float cy, float near, float far, const int32_t* step, int64_t n_rand,
                          float* rays, float* target, int32_t* pix, void* stream);

/* ---- section 8f "next" row 2: total_variation_loss : loss.py:11-43 ---------------------------------- */
/* img2mse (run_nerf_helpers.py:24): out[0] = mean((a - b)^2) over n floats, one CTA, fixed summation order.
 * Backward: da = gout[0] * 2 (a - b) / n, db = -da; either of da / db may be NULL. */
HN_API int hn_mse_fwd(const float* a, const float* b, int64_t n, float* out, void* stream);
HN_API int hn_mse_bwd(const float* a, const float* b, int64_t n, const float* gout, float* da, float* db,
                      void* stream);

/* One hash level: table [2^log2T, F]; origin = int64[3] on the device (the random cube corner drawn at
 * loss.py:25); cube = cube size (loss.py:22).  fwd writes out[0] = (sum of squared forward differences over the
 * (cube+1)^3 hashed vertices) / cube.  bwd ACCUMULATES gout[0] * d out / d table into dtable [2^log2T, F]. */
HN_API int hn_tv_loss_fwd(const float* table, const int64_t* origin, int cube, int log2T, int F, float* out,
                          void* stream);
HN_API int hn_tv_loss_bwd(const float* table, const int64_t* origin, int cube, int log2T, int F, const float* gout,
                          float* dtable, void* stream);
/* All L levels of a sweep (run_nerf.py:628-635 calls total_variation_loss once per level) in one launch:
 * tables [L, 2^log2T, F] flat, origins int64 [L,3] and cubes int32 [L] on the device, max_cube = max(cubes)
 * (sizes the grid), out / gout [L], dtables accumulated like hn_tv_loss_bwd, level by level. */
HN_API int hn_tv_loss_fwd_levels(const float* tables, const int64_t* origins, const int32_t* cubes, int L,
                                 int max_cube, int log2T, int F, float* out, void* stream);
HN_API int hn_tv_loss_bwd_levels(const float* tables, const int64_t* origins, const int32_t* cubes, int L,
                                 int max_cube, int log2T, int F, const float* gout, float* dtables, void* stream);

/* CUDA-graph friendly form: the step-dependent scalars come from device memory,
 * hp = {beta1, beta2, eps, weight_decay*lr, step_size*lr, grad_scale, mode, zero_grad}. */
HN_API int hn_radam_step_dev(float* p, float* g, float* m, float* v, int64_t n, const float* hp, void* stream);

/* ---- section 8e: the data-parallel exchange over peer memory ------------------------------------------------ */
/* One rank per GPU.  All buffers are SYMMETRIC allocations (same size / layout on every rank) whose peer mappings
 * the caller sets up (torch.distributed._symmetric_memory in hn_b200/dp.py); `*_ptrs_dev` are DEVICE arrays of the
 * `world` peer pointers, `*_mc` multicast pointers to the same buffers (NVSwitch / NVLS) or NULL.
 *
 * hn_dp_barrier: every rank writes `epoch` into its flag of slot `slot` (0..3) in each peer's signal buffer
 * (>= 256 uint32 per rank, zero-initialised; epochs must grow by one per use of a slot) and waits for all peers'
 * flags; release / acquire at system scope.  A wait longer than ~2 s traps (a peer died) instead of hanging.
 *
 * hn_dp_reduce_update: over a span of n floats (multiple of 4; pointers 16-byte aligned) this rank takes slice
 * rank of ceil(n / world) and, in one pass: sums the slice of every rank's gradient (multimem.ld_reduce when
 * grad_mc != NULL, else one peer load per rank), applies the RAdam update of hn_radam_step_dev to ITS slice of
 * (params, m, v) with hp = {beta1, beta2, eps, weight_decay*lr, step_size*lr, grad_scale, mode, -} read from device
 * memory, stores the new parameters to every rank (multimem.st / peer stores) and clears the slice of every rank's
 * gradient.  mode = -1: plain summed all-reduce in place of the gradient buffers, no optimizer (params, m, v unused).
 * Bracket the calls of one exchange with two hn_dp_barrier calls (different slots). */
HN_API int hn_dp_barrier(void* const* signal_ptrs_dev, int rank, int world, int slot, uint32_t epoch, void* stream);
HN_API int hn_dp_reduce_update(void* const* grad_ptrs_dev, float* grad_mc, void* const* param_ptrs_dev, float* param_mc,
                               float* m, float* v, int rank, int world, int64_t n, const float* hp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HASHNERF_B200_H_ */
