"""CPU oracle for the HashNeRF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain fp32 torch-on-CPU / numpy, of the arithmetic the reference performs on
the path ``HashEmbedder -> SHEncoder -> NeRFSmall -> raw2outputs / sample_pdf`` (SURVEY.md
section 8a, Appendix A).  Every function cites the reference lines it restates; paths are
relative to the reference checkout (mache102/HashNeRF-pytorch).

Rules (task tier section 3):
* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
  leg may import this module, and only as the checker or the timed CPU baseline;
* the product package (``hashnerf-pytorch_b200/``) never imports it and has no CPU fallback.

Pinning: the reference ships no tests, golden vectors or fixtures for this path (SURVEY 8c:
"parity unpinned" by the reference itself).  This oracle is therefore pinned against OUTPUTS OF
THE REFERENCE ITSELF: ``oracle/gen_golden.py`` executes the unmodified reference functions
(through ``oracle/ref_loader.py``) in the build container and commits the vectors under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below against them
(bit-exact for indices and for the fp32 forward arithmetic, which uses the same ATen CPU ops in
the same order).

Why torch-on-CPU rather than numpy/C: the reference *is* fp32 ATen arithmetic; using the same
correctly-rounded elementwise primitives in the reference's order makes the oracle bit-exact on
the forward path, and autograd over the restated forward is the backward oracle.  Integer work
(the spatial hash) is additionally restated in numpy (``spatial_hash_np``).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------
# constants
# ----------------------------------------------------------------------------------------------
# embedding/hash_encoding.py:7 -- per-dimension multipliers of the Teschner spatial hash.
PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)

# embedding/hash_encoding.py:10 -- corner c = 4*i + 2*j + k, (i, j, k) offsets along (x, y, z).
CORNER_OFFSETS = np.array([[(c >> 2) & 1, (c >> 1) & 1, c & 1] for c in range(8)], dtype=np.int64)

# embedding/spherical_harmonic.py:13-30 -- real SH constants, degrees 0..3.
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435)
# embedding/spherical_harmonic.py:31-41 -- degree 4 (only used when degree == 5).
SH_C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892,
         0.10578554691520431, -0.6690465435572892, 0.47308734787878004, -1.7701307697799304,
         0.6258357354491761)


# ----------------------------------------------------------------------------------------------
# (a1) level resolutions
# ----------------------------------------------------------------------------------------------
def growth_factor(base_resolution: int, finest_resolution: int, n_levels: int) -> torch.Tensor:
    """embedding/hash_encoding.py:46-50 -- b = exp((ln finest - ln base)/(L-1)), fp32 0-dim.

    The reference takes ``log`` of int64 0-dim tensors (promoted to fp32)."""
    lo = torch.tensor(base_resolution, device="cpu")
    hi = torch.tensor(finest_resolution, device="cpu")
    return torch.exp((torch.log(hi) - torch.log(lo)) / (n_levels - 1))


def level_resolutions(base_resolution: int = 16, finest_resolution: int = 512,
                      n_levels: int = 16) -> torch.Tensor:
    """embedding/hash_encoding.py:101 -- floor(base * b**i) per level, as an fp32 [L] tensor."""
    b = growth_factor(base_resolution, finest_resolution, n_levels)
    base = torch.tensor(base_resolution, device="cpu")
    return torch.stack([torch.floor(base * b ** i) for i in range(n_levels)]).to(torch.float32)


def tv_level_resolution(min_resolution, max_resolution, level: int, n_levels: int = 16) -> int:
    """loss.py:13-14 -- the TV loss recomputes the resolution in fp64 ``math`` (Appendix B10)."""
    b = math.exp((math.log(max_resolution) - math.log(min_resolution)) / (n_levels - 1))
    return int(math.floor(min_resolution * b ** level))


# ----------------------------------------------------------------------------------------------
# (a3) spatial hash
# ----------------------------------------------------------------------------------------------
def spatial_hash(coords: torch.Tensor, log2_hashmap_size: int) -> torch.Tensor:
    """embedding/hash_encoding.py:112-128 -- XOR over dims of coord*prime, masked to T bits (int64)."""
    c = coords.to(torch.int64)
    acc = torch.zeros(c.shape[:-1], dtype=torch.int64, device=c.device)
    for d in range(c.shape[-1]):
        acc = acc ^ (c[..., d] * PRIMES[d])
    return acc & ((1 << log2_hashmap_size) - 1)


def spatial_hash_np(coords: np.ndarray, log2_hashmap_size: int) -> np.ndarray:
    """numpy uint32 restatement: wrap-around arithmetic keeps the low T<=32 bits identical."""
    c = np.asarray(coords).astype(np.int64).astype(np.uint32)
    acc = np.zeros(c.shape[:-1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for d in range(c.shape[-1]):
            acc ^= c[..., d] * np.uint32(PRIMES[d] & 0xFFFFFFFF)
    mask = np.uint32(((1 << log2_hashmap_size) - 1) & 0xFFFFFFFF)
    return (acc & mask).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# (a2) voxel vertices of one level
# ----------------------------------------------------------------------------------------------
def voxel_vertices(xyz: torch.Tensor, box_min: torch.Tensor, box_max: torch.Tensor,
                   resolution: torch.Tensor, log2_hashmap_size: int):
    """embedding/hash_encoding.py:59-82 for one level.

    Returns (xyz_after_clamp, cell_idx int32 [N,3], vmin, vmax, hashed int64 [N,8], inside [N,3]).
    The clamp is applied only if some point is outside (:67-69); clamping in-box points is the
    identity so the branch is arithmetic-neutral and is restated unconditionally."""
    inside = xyz == torch.max(torch.min(xyz, box_max), box_min)                       # :66
    xyz = torch.max(torch.min(xyz, box_max), box_min)                                 # :69
    cell = (box_max - box_min) / resolution                                           # :72
    idx = torch.floor((xyz - box_min) / cell).to(torch.int32)                         # :74
    vmin = idx * cell + box_min                                                       # :75 (mul, add)
    vmax = vmin + cell                                                                # :76 (1.0*cell exact)
    corners = idx.to(torch.int64)[:, None, :] + torch.from_numpy(CORNER_OFFSETS)[None]  # :79
    hashed = spatial_hash(corners, log2_hashmap_size)                                 # :80
    return xyz, idx, vmin, vmax, hashed, inside


def trilerp(x: torch.Tensor, vmin: torch.Tensor, vmax: torch.Tensor, corner_feats: torch.Tensor):
    """embedding/hash_encoding.py:130-163 -- x-lerp of corner pairs (c, c+4), then y, then z.

    Each ``a*(1-w) + b*w`` is four separately rounded fp32 ops (sub, mul, mul, add)."""
    w = (x - vmin) / (vmax - vmin)                                                    # :143
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    e = corner_feats
    along_x = [e[:, c] * (1 - wx) + e[:, c + 4] * wx for c in range(4)]               # :149-152
    along_y = [along_x[k] * (1 - wy) + along_x[k + 2] * wy for k in range(2)]         # :156-157
    return along_y[0] * (1 - wz) + along_y[1] * wz                                    # :161


def corner_weights(x, vmin, vmax):
    """The eight scalar weights the lerp chain is algebraically equal to (used for the backward
    formula dtable[h_c] += dy * W_c, SURVEY A.2 last paragraph)."""
    w = (x - vmin) / (vmax - vmin)
    out = []
    for c in range(8):
        f = torch.ones_like(w[:, 0])
        for axis, bit in enumerate(((c >> 2) & 1, (c >> 1) & 1, c & 1)):
            f = f * (w[:, axis] if bit else (1 - w[:, axis]))
        out.append(f)
    return torch.stack(out, dim=1)  # [N, 8]


# ----------------------------------------------------------------------------------------------
# (a6) full encoder
# ----------------------------------------------------------------------------------------------
def hash_encode(x: torch.Tensor, tables: Sequence[torch.Tensor] | torch.Tensor,
                box_min: torch.Tensor, box_max: torch.Tensor, resolutions: torch.Tensor,
                log2_hashmap_size: int, return_debug: bool = False):
    """embedding/hash_encoding.py:84-110.

    ``tables``: [L, 2^T, F] tensor or a list of L [2^T, F] tensors.  Quirks kept:
    the clamped coordinates persist across levels (``self.xyz`` is overwritten, :69), the lerp
    weights use the UNCLAMPED input (:107), and the returned mask is the last level's (:109)."""
    running = x
    feats, dbg = [], []
    inside = None
    for lvl in range(len(resolutions)):
        running, idx, vmin, vmax, hashed, inside = voxel_vertices(
            running, box_min, box_max, resolutions[lvl], log2_hashmap_size)
        corner_feats = tables[lvl][hashed]                                            # :106 gather
        feats.append(trilerp(x, vmin, vmax, corner_feats))                            # :107
        if return_debug:
            dbg.append(dict(idx=idx, vmin=vmin, vmax=vmax, hashed=hashed))
    keep = inside.sum(dim=-1) == inside.shape[-1]                                     # :109
    out = torch.cat(feats, dim=-1)                                                    # :110
    return (out, keep, dbg) if return_debug else (out, keep)


def hash_encode_grad_tables(x, dy, box_min, box_max, resolutions, log2_hashmap_size,
                            n_features: int, dtype=torch.float64) -> torch.Tensor:
    """Analytic table gradient (what autograd of :106-:161 yields), accumulated in ``dtype``.

    dtable_l[h_c] += dy[:, l*F:(l+1)*F] * W_c.  fp64 accumulation gives an order-independent
    yardstick for the atomically accumulated CUDA result."""
    L = len(resolutions)
    T = 1 << log2_hashmap_size
    out = torch.zeros(L, T, n_features, dtype=dtype)
    running = x
    for lvl in range(L):
        running, idx, vmin, vmax, hashed, _ = voxel_vertices(
            running, box_min, box_max, resolutions[lvl], log2_hashmap_size)
        W = corner_weights(x, vmin, vmax).to(dtype)                                   # [N, 8]
        g = dy[:, lvl * n_features:(lvl + 1) * n_features].to(dtype)                  # [N, F]
        contrib = W[:, :, None] * g[:, None, :]                                       # [N, 8, F]
        out[lvl].index_add_(0, hashed.reshape(-1), contrib.reshape(-1, n_features))
    return out


# ----------------------------------------------------------------------------------------------
# (a7) spherical harmonics
# ----------------------------------------------------------------------------------------------
def sh_encode(dirs: torch.Tensor, degree: int = 4) -> torch.Tensor:
    """embedding/spherical_harmonic.py:65-103.  Python-float constants multiply fp32 tensors, so
    each ``C * t`` rounds C to fp32 at the multiply; products are evaluated left to right."""
    assert dirs.shape[-1] == 3 and 1 <= degree <= 5                                   # :51-52
    x, y, z = dirs.unbind(-1)
    cols = [torch.full_like(x, SH_C0)]                                                # :70
    if degree > 1:
        cols += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]                                   # :72-74
    if degree > 2:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        cols += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.0 * zz - xx - yy),
                 SH_C2[3] * xz, SH_C2[4] * (xx - yy)]                                 # :78-83
    if degree > 3:
        cols += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z,
                 SH_C3[2] * y * (4 * zz - xx - yy), SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
                 SH_C3[4] * x * (4 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
                 SH_C3[6] * x * (xx - 3 * yy)]                                        # :85-91
    if degree > 4:
        cols += [SH_C4[0] * xy * (xx - yy), SH_C4[1] * yz * (3 * xx - yy),
                 SH_C4[2] * xy * (7 * zz - 1), SH_C4[3] * yz * (7 * zz - 3),
                 SH_C4[4] * (zz * (35 * zz - 30) + 3), SH_C4[5] * xz * (7 * zz - 3),
                 SH_C4[6] * (xx - yy) * (7 * zz - 1), SH_C4[7] * xz * (xx - 3 * yy),
                 SH_C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]                # :93-101
    return torch.stack(cols, dim=-1)


# ----------------------------------------------------------------------------------------------
# (a8) NeRFSmall
# ----------------------------------------------------------------------------------------------
def nerf_small(x: torch.Tensor, sigma_w: Sequence[torch.Tensor], color_w: Sequence[torch.Tensor],
               input_ch: int = 32) -> torch.Tensor:
    """models.py:151-174 -- bias-free sigma net then colour net; out = [rgb_raw(3) | sigma(1)].

    ``sigma_w`` / ``color_w`` are the ``nn.Linear.weight`` matrices ([out, in]) in layer order."""
    h = x[..., :input_ch]
    views = x[..., input_ch:]
    for i, w in enumerate(sigma_w):                                                   # :155-159
        h = h @ w.t()
        if i != len(sigma_w) - 1:
            h = torch.relu(h)
    sigma, geo = h[..., 0], h[..., 1:]                                                # :161
    h = torch.cat([views, geo], dim=-1)                                               # :164 (SH first)
    for i, w in enumerate(color_w):                                                   # :165-168
        h = h @ w.t()
        if i != len(color_w) - 1:
            h = torch.relu(h)
    return torch.cat([h, sigma[..., None]], dim=-1)                                   # :172


def run_network(pts: torch.Tensor, viewdirs: Optional[torch.Tensor], mlp_weights, enc, sh_degree=4):
    """run_nerf_helpers.py:212-227.  ``enc(points[N,3]) -> (features, keep)``; ``mlp_weights`` is
    ``(sigma_w, color_w)``.  Sigma is zeroed where ``~keep`` (:225)."""
    flat = pts.reshape(-1, pts.shape[-1])
    feats, keep = enc(flat)
    if viewdirs is not None:
        d = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)                        # :219-220
        feats = torch.cat([feats, sh_encode(d, sh_degree)], dim=-1)                   # :221-222
    out = nerf_small(feats, mlp_weights[0], mlp_weights[1], input_ch=feats.shape[-1] - (
        0 if viewdirs is None else sh_degree ** 2))
    sig = torch.where(keep, out[:, -1], torch.zeros_like(out[:, -1]))                 # :225
    out = torch.cat([out[:, :-1], sig[:, None]], dim=-1)
    return out.reshape(*pts.shape[:-1], out.shape[-1])


# ----------------------------------------------------------------------------------------------
# (a10) compositing
# ----------------------------------------------------------------------------------------------
_EPS32 = float(torch.finfo(torch.float32).eps)


def composite(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor,
              noise: Optional[torch.Tensor] = None, white_bkgd: bool = False):
    """run_nerf_helpers.py:577-628.  ``noise`` is the already-scaled additive sigma noise
    (``randn * raw_noise_std``, :599-606) or None.

    Returns (rgb_map, disp_map, acc_map, weights, depth_map, entropy)."""
    R, S = z_vals.shape
    gaps = z_vals[:, 1:] - z_vals[:, :-1]                                             # :592
    gaps = torch.cat([gaps, torch.full((R, 1), 1e10, dtype=z_vals.dtype)], dim=-1)    # :593
    gaps = gaps * torch.norm(rays_d[:, None, :], dim=-1)                              # :595
    rgb = torch.sigmoid(raw[..., :3])                                                 # :597
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise                     # :609
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * gaps)                                # :590
    trans = torch.cumprod(torch.cat([torch.ones(R, 1), 1.0 - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    weights = alpha * trans                                                           # :611
    rgb_map = torch.sum(weights[..., None] * rgb, dim=-2)                             # :612
    acc_map = torch.sum(weights, dim=-1)                                              # :616
    depth_map = torch.sum(weights * z_vals, dim=-1) / acc_map                         # :614
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map)         # :615
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])                                # :618-619
    # :623 Categorical(probs=[w, 1 - sum(w) + 1e-6]).entropy(): probs are normalised by their sum,
    # logits = log(clamp(p, eps, 1-eps)), entropy = -sum(p * logits)  (torch/distributions).
    q = torch.cat([weights, 1.0 - weights.sum(-1, keepdim=True) + 1e-6], dim=-1)
    p = q / q.sum(-1, keepdim=True)
    logit = torch.log(p.clamp(min=_EPS32, max=1.0 - _EPS32))
    entropy = -(logit * p).sum(-1)
    return rgb_map, disp_map, acc_map, weights, depth_map, entropy


# ----------------------------------------------------------------------------------------------
# (a11) hierarchical sampling, (a12) ray marching set-up
# ----------------------------------------------------------------------------------------------
def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """run_nerf_helpers.py:264-307 with the uniform variates ``u`` [R, Ni] passed in (the
    reference draws them at :274-276 or builds the linspace at :271-272)."""
    w = weights + 1e-5                                                                # :266
    pdf = w / torch.sum(w, -1, keepdim=True)                                          # :267
    cdf = torch.cumsum(pdf, -1)                                                       # :268
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)                        # :269
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                                     # :290
    below = (inds - 1).clamp(min=0)                                                   # :291
    above = inds.clamp(max=cdf.shape[-1] - 1)                                         # :292
    c_lo, c_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)             # :298
    b_lo, b_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)           # :299
    denom = c_hi - c_lo                                                               # :301
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)                  # :302
    t = (u - c_lo) / denom                                                            # :303
    return b_lo + t * (b_hi - b_lo)                                                   # :304


def sample_pdf_tolerance(bins: torch.Tensor, weights: torch.Tensor, u: torch.Tensor, rtol: float = 1e-5,
                         atol: float = 2e-5, cdf_eps: float = 1e-6):
    """Deterministic per-sample acceptance bound for an implementation of run_nerf_helpers.py:264-307 whose CDF is
    summed in another order than ATen's sequential cumsum (ours is a warp scan).

    The CDF entries of two correct fp32 implementations differ by rounding noise |d| <= ``cdf_eps`` (a 64..192-term
    fp32 prefix sum of values <= 1: worst case ~ n * 2^-24 ~ 4e-6..1e-5, observed ~ 1e-7; 1e-6 = 16 ulp(1) is the bar
    we hold ourselves to).  Forward error analysis of :303-304, ``t = (u - c_lo) / denom``,
    ``sample = b_lo + t (b_hi - b_lo)``:
      * regular samples: t moves by <= 2 cdf_eps / denom, the sample by that times the bin width;
      * BRANCH-CHAOTIC samples: ``denom`` within 2 cdf_eps of the ``denom < 1e-5 -> 1`` switch (:302) -- which branch
        runs is decided by rounding noise in the reference itself (CPU and CUDA ATen disagree) -- or ``u`` within
        cdf_eps of a CDF entry (searchsorted may pick the neighbouring bin): anywhere inside the two bins involved.
    Returns (want, tol, chaotic) with |got - want| <= tol required for EVERY sample."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    last = cdf.shape[-1] - 1
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=last)
    c_lo, c_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b_lo, b_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom_raw = c_hi - c_lo
    denom = torch.where(denom_raw < 1e-5, torch.ones_like(denom_raw), denom_raw)
    want = b_lo + (u - c_lo) / denom * (b_hi - b_lo)
    width = (b_hi - b_lo).abs()
    # widths of the neighbouring bins, for samples that may land next door
    w_prev = (b_lo - torch.gather(bins, 1, (below - 1).clamp(min=0))).abs()
    w_next = (torch.gather(bins, 1, (above + 1).clamp(max=last)) - b_hi).abs()
    chaotic = ((denom_raw - 1e-5).abs() <= 2 * cdf_eps) | ((u - c_lo).abs() <= cdf_eps) | ((c_hi - u).abs() <= cdf_eps)
    regular = width * torch.clamp(2 * cdf_eps / denom, max=1.0)
    tol = rtol * want.abs() + atol + torch.where(chaotic, width + torch.maximum(w_prev, w_next), regular)
    return want, tol, chaotic


def det_u(n_rays: int, n_importance: int) -> torch.Tensor:
    """run_nerf_helpers.py:271-272."""
    return torch.linspace(0.0, 1.0, steps=n_importance).expand(n_rays, n_importance)


def coarse_z(near: torch.Tensor, far: torch.Tensor, n_samples: int, lindisp: bool,
             t_rand: Optional[torch.Tensor]) -> torch.Tensor:
    """run_nerf_helpers.py:514-536.  ``near``/``far`` are [R,1]; ``t_rand`` [R,S] in [0,1) or None."""
    t = torch.linspace(0.0, 1.0, steps=n_samples)                                     # :514
    if not lindisp:
        z = near * (1.0 - t) + far * t                                                # :516
    else:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)                            # :518
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mids = 0.5 * (z[:, 1:] + z[:, :-1])                                           # :524
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * t_rand                                          # :536
    return z


def render_rays(ray_batch: torch.Tensor, enc, coarse_w, fine_w, n_samples: int, n_importance: int,
                t_rand=None, u=None, noise0=None, noise1=None, lindisp=False, white_bkgd=False,
                perturb: float = 0.0, sh_degree: int = 4, z_samples=None):
    """run_nerf_helpers.py:464-574 with every random draw passed in explicitly.

    ``ray_batch`` [R, 11] = (o, d, near, far, viewdir).  Returns the reference's output dict plus the inputs and
    the output of the resampling step (``pdf_bins``, ``pdf_weights``, ``pdf_u``, ``z_samples``).
    ``z_samples`` [R, Ni], if given, replaces the output of sample_pdf (:548): a checker can hand in the depths
    the implementation under test drew -- after holding them to ``sample_pdf_tolerance`` -- so that everything
    downstream is compared on identical sample positions with the strict tolerance."""
    R = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None                 # :510
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z = coarse_z(near, far, n_samples, lindisp, t_rand if perturb > 0 else None)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]                     # :538
    raw = run_network(pts, viewdirs, coarse_w, enc, sh_degree)
    rgb, disp, acc, wts, depth, ent = composite(raw, z, rays_d, noise0, white_bkgd)   # :541
    ret = {}
    if n_importance > 0:
        rgb0, depth0, acc0, ent0 = rgb, depth, acc, ent
        mids = 0.5 * (z[:, 1:] + z[:, :-1])                                           # :547
        if u is None:
            u = det_u(R, n_importance)
        z_new = sample_pdf(mids, wts[:, 1:-1], u).detach()                            # :548-549
        ret.update(pdf_bins=mids.detach(), pdf_weights=wts[:, 1:-1].detach(), pdf_u=u, z_samples_own=z_new)
        if z_samples is not None:
            z_new = z_samples.detach()
        ret.update(z_samples=z_new)
        z, _ = torch.sort(torch.cat([z, z_new], -1), -1)                              # :551
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]                 # :552
        raw = run_network(pts, viewdirs, fine_w if fine_w is not None else coarse_w, enc, sh_degree)
        rgb, disp, acc, wts, depth, ent = composite(raw, z, rays_d, noise1, white_bkgd)
        ret.update(rgb0=rgb0, depth0=depth0, acc0=acc0, sparsity_loss0=ent0,
                   z_std=torch.std(z_new, dim=-1, unbiased=False))                    # :563-568
    ret.update(rgb_map=rgb, depth_map=depth, acc_map=acc, sparsity_loss=ent, raw=raw,
               weights=wts, z_vals=z, disp_map=disp)
    return ret


# ----------------------------------------------------------------------------------------------
# section 8f "next" row 3: ray generation
# ----------------------------------------------------------------------------------------------
def pinhole_rays(H: int, W: int, K, c2w: torch.Tensor):
    """ray_util.py:62-80 -- (rays_o, rays_d), each [H, W, 3]."""
    cols = torch.linspace(0, W - 1, W)
    rows = torch.linspace(0, H - 1, H)
    j, i = torch.meshgrid(rows, cols, indexing="ij")                                  # :71-73 (ij meshgrid, then .t())
    cam = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)   # :74
    rays_d = torch.sum(cam[..., None, :] * c2w[:3, :3], -1)                           # :77
    rays_o = c2w[:3, -1].expand(rays_d.shape)                                         # :79
    return rays_o, rays_d


def ndc_rays(H: int, W: int, focal: float, near: float, rays_o: torch.Tensor, rays_d: torch.Tensor):
    """ray_util.py:96-142 -- forward-facing scenes: shift origins to the near plane, project to NDC."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]                                     # :119
    rays_o = rays_o + t[..., None] * rays_d                                           # :120
    ox_oz = rays_o[..., 0] / rays_o[..., 2]                                           # :124
    oy_oz = rays_o[..., 1] / rays_o[..., 2]
    o0 = -1. / (W / (2. * focal)) * ox_oz                                             # :129
    o1 = -1. / (H / (2. * focal)) * oy_oz
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - ox_oz)         # :134
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - oy_oz)
    d2 = 1 - o2
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)               # :139-140


# ----------------------------------------------------------------------------------------------
# section 8f "next" rows: TV loss and RAdam (restated for the kernels that replace them)
# ----------------------------------------------------------------------------------------------
def total_variation(table: torch.Tensor, min_res: int, max_res: int, level: int,
                    log2_hashmap_size: int, n_levels: int, min_vertex: torch.Tensor):
    """loss.py:11-43 with the random cube origin ``min_vertex`` [3] (drawn at :25) passed in."""
    res = tv_level_resolution(min_res, max_res, level, n_levels)                      # :13-14
    cube = int(math.floor(min(max(res / 10.0, min_res - 1), 50)))                     # :17-22
    ax = torch.arange(cube + 1)
    gx, gy, gz = torch.meshgrid(min_vertex[0] + ax, min_vertex[1] + ax, min_vertex[2] + ax,
                                indexing="ij")                                        # :26-27
    e = table[spatial_hash(torch.stack([gx, gy, gz], -1), log2_hashmap_size)]         # :29-30
    tv = ((e[1:] - e[:-1]) ** 2).sum() + ((e[:, 1:] - e[:, :-1]) ** 2).sum() \
        + ((e[:, :, 1:] - e[:, :, :-1]) ** 2).sum()                                   # :39-41
    return tv / cube                                                                  # :43


def tv_cube_size(min_res: int, max_res: int, level: int, n_levels: int = 16) -> Tuple[int, int]:
    """(resolution, cube_size) of loss.py:13-22."""
    res = tv_level_resolution(min_res, max_res, level, n_levels)
    return res, int(math.floor(min(max(res / 10.0, min_res - 1), 50)))


def radam_step(p, g, m, v, step: int, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float):
    """radam.py:34-92 for one tensor, in place on (p, m, v); ``step`` is the 1-based count."""
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)                                     # :58
    m.mul_(beta1).add_(g, alpha=1 - beta1)                                            # :59
    beta2_t = beta2 ** step
    n_max = 2 / (1 - beta2) - 1
    n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)                                # :66-68
    if n_sma >= 5:
        step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma
                              * n_max / (n_max - 2)) / (1 - beta1 ** step)            # :73
        if weight_decay != 0:
            p.add_(p, alpha=-weight_decay * lr)                                       # :82-83
        p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size * lr)                      # :84-85
    # else: step_size = -1 (degenerated_to_sgd False) -> parameters untouched          :76-77,87
    return p
