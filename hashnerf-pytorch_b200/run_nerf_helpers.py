"""Render orchestration -- drop-in for the reference's ``run_nerf_helpers.py``.

Public names and signatures follow the reference (create_nerf :51, batchify :203, run_network :212,
get_embedder :230, sample_pdf :264, render :310, render_path :395, render_rays :464, raw2outputs :577,
img2mse / mse2psnr / to8b :24-26, device :28) so that an unmodified ``run_nerf.py`` can
``from run_nerf_helpers import *``.  What differs is what executes: every per-sample stage is a CUDA
kernel from libhashnerf_b200.so (see hn_b200/ops.py), the per-sample expansion of view directions and the
[N,48] concatenation never exist, and none of the reference's ~70 host synchronisations per step
(torch.all in the encoder, boolean-mask index_put_, Categorical validation, isnan checks) remain.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (star-exported like the reference)

from hn_b200 import autograph, ops
from radam import RAdam
from ray_util import get_rays, get_rays_np, get_ndc_rays  # noqa: F401
from embedding.embedder import Embedder
from embedding.hash_encoding import HashEmbedder, SHEncoder
from models import NeRF, NeRFSmall, NeRFGradient  # noqa: F401

# Misc (run_nerf_helpers.py:24-26)
def img2mse(x, y):
    """mean((x - y) ** 2) (run_nerf_helpers.py:24).  Same-shape fp32 CUDA tensors: one launch each way (hn_mse_fwd /
    hn_mse_bwd) instead of three forward and four backward; anything else: the reference's expression."""
    if (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor) and x.is_cuda and y.is_cuda
            and x.dtype == torch.float32 and y.dtype == torch.float32 and x.shape == y.shape and x.numel() > 0):
        return ops.mse(x, y)
    return torch.mean((x - y) ** 2)
# the reference's torch.Tensor([10.]) lands on the default tensor type's device (CUDA after run_nerf.py:725); built
# on x's device here so that the helper also works without that global switch -- same value, same [1] shape
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], dtype=torch.float32, device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
DEBUG = False
# create_nerf's RAdam clears the gradients it consumed in its own pass over memory, so that the
# ``optimizer.zero_grad()`` of the next iteration (run_nerf.py:612) has no 64 MiB buffer left to fill.  Invisible to
# run_nerf.py (it never reads .grad after step()); set to False for gradients that survive step().
FUSED_ZERO_GRAD = True


# ----------------------------------------------------------------------------------------------
# network evaluation
# ----------------------------------------------------------------------------------------------
def batchify(fn, chunk):
    """Apply ``fn`` in row chunks (run_nerf_helpers.py:203-210).  Kept for API parity; the fused MLP does
    not need it because it materialises no per-sample intermediates."""
    if chunk is None:
        return fn

    def chunked(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return chunked


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """Encode points (and view directions) and evaluate ``fn`` (run_nerf_helpers.py:212-227).

    inputs [R,S,3], viewdirs [R,3] -> [R,S,4].  With our modules the SH features are evaluated once per
    ray and handed to the fused MLP together with the keep mask; any other combination of callables goes
    through the generic expand / cat / mask sequence of the reference."""
    flat = inputs.reshape(-1, inputs.shape[-1])
    per_ray = inputs.shape[-2] if inputs.dim() >= 3 else 1

    if isinstance(fn, NeRFSmall) and fn.fused and isinstance(embed_fn, HashEmbedder) and viewdirs is not None \
            and isinstance(embeddirs_fn, SHEncoder) and embeddirs_fn.degree == 4 \
            and viewdirs.shape[0] * per_ray == flat.shape[0]:
        embedded, keep_u8 = embed_fn.encode(flat, ordered=inputs.dim() >= 3)  # [R,S,3]: samples along rays
        out = fn.forward_fused(embedded, embeddirs_fn(viewdirs), per_ray, keep_u8)
        return out.reshape(*inputs.shape[:-1], 4)

    embedded, keep_mask = embed_fn(flat)
    if viewdirs is not None:
        dirs = viewdirs[:, None].expand(inputs.shape).reshape(-1, inputs.shape[-1])
        embedded = torch.cat([embedded, embeddirs_fn(dirs)], -1)
    out = batchify(fn, netchunk)(embedded)
    out = torch.cat([out[..., :-1], torch.where(keep_mask, out[..., -1], torch.zeros_like(out[..., -1]))[..., None]], -1)
    return out.reshape(*inputs.shape[:-1], out.shape[-1])


def get_embedder(multires, args, i=0):
    """-1: identity, 0: frequency encoding, 1: hash encoding, 2: spherical harmonics (:230-260)."""
    if i == -1:
        return nn.Identity(), 3
    if i == 0:
        enc = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                       log_sampling=True, periodic_fns=[torch.sin, torch.cos])
        return (lambda x, eo=enc: eo.embed(x)), enc.out_dim
    if i == 1:
        enc = HashEmbedder(bounding_box=args.bounding_box, log2_hashmap_size=args.log2_hashmap_size,
                           finest_resolution=args.finest_res)
        return enc, enc.out_dim
    if i == 2:
        enc = SHEncoder()
        return enc, enc.out_dim
    raise ValueError(f"unknown embedder id {i}")


# ----------------------------------------------------------------------------------------------
# model / optimizer factory (run_nerf_helpers.py:51-200)
# ----------------------------------------------------------------------------------------------
def _make_network(args, input_ch, input_ch_views, fine):
    if args.i_embed == 1:
        return NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                         input_ch=input_ch, input_ch_views=input_ch_views).to(device)
    cls = NeRFGradient if getattr(args, "use_gradient", False) else NeRF
    return cls(D=args.netdepth_fine if fine else args.netdepth, W=args.netwidth_fine if fine else args.netwidth,
               input_ch=input_ch, output_ch=5 if args.N_importance > 0 else 4, skips=[4],
               input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(device)


def create_nerf(args):
    """Instantiate encoders, coarse/fine networks and the optimizer; reload the newest checkpoint.

    Returns (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer)."""
    if autograph.ENABLED and torch.cuda.is_available() and torch.device(device).type == "cuda":
        autograph.ensure_stream(device)   # HN_AUTO_GRAPH=1: everything from here on runs on one non-default stream
    embed_fn, input_ch = get_embedder(args.multires, args, i=args.i_embed)
    embedding_params = list(embed_fn.parameters()) if args.i_embed == 1 else []
    if isinstance(embed_fn, nn.Module):
        embed_fn.to(device)

    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args, i=args.i_embed_views)

    model = _make_network(args, input_ch, input_ch_views, fine=False)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = _make_network(args, input_ch, input_ch_views, fine=True)
        grad_vars += list(model_fine.parameters())

    def network_query_fn(inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
                           netchunk=args.netchunk)

    if args.i_embed == 1:
        optimizer = RAdam([{'params': grad_vars, 'weight_decay': 1e-6},
                           {'params': embedding_params, 'eps': 1e-15}], lr=args.lrate, betas=(0.9, 0.99),
                          fused_zero_grad=FUSED_ZERO_GRAD)
    else:
        optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        run_dir = os.path.join(args.basedir, args.expname)
        ckpts = [os.path.join(run_dir, f) for f in sorted(os.listdir(run_dir)) if 'tar' in f]
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        print('Reloading from', ckpts[-1])
        ckpt = torch.load(ckpts[-1], map_location=device, weights_only=False)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])
        if args.i_embed == 1:
            embed_fn.load_state_dict(ckpt['embed_fn_state_dict'])

    render_kwargs_train = {
        'network_query_fn': network_query_fn,
        'perturb': args.perturb,
        'N_importance': args.N_importance,
        'network_fine': model_fine,
        'N_samples': args.N_samples,
        'network_fn': model,
        'embed_fn': embed_fn,
        'use_viewdirs': args.use_viewdirs,
        'white_bkgd': args.white_bkgd,
        'raw_noise_std': args.raw_noise_std,
    }
    if (args.dataset_type not in ['llff', 'st3d']) or args.no_ndc:
        print('Not ndc!')
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


# ----------------------------------------------------------------------------------------------
# hierarchical sampling / compositing
# ----------------------------------------------------------------------------------------------
_LINSPACE = {}


def _linspace01(steps, dev):
    """torch.linspace(0, 1, steps) on ``dev``, cached (it is recomputed by the reference on every call)."""
    key = (int(steps), dev.type, dev.index)
    t = _LINSPACE.get(key)
    if t is None:
        t = _LINSPACE[key] = torch.linspace(0., 1., steps=steps, device=dev)
    return t


def _uniform_variates(R, n, det, pytest, dev):
    """(u [R,n] or None, u_det [n] or None) exactly as sample_pdf draws them (run_nerf_helpers.py:270-287)."""
    if pytest:  # fixed numpy variates
        np.random.seed(0)
        if det:
            u = np.broadcast_to(np.linspace(0., 1., n), (R, n))
        else:
            u = np.random.rand(R, n)
        return torch.Tensor(np.ascontiguousarray(u)).to(dev), None
    if det:
        return None, _linspace01(n, dev)
    return torch.rand(R, n, device=dev), None


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """Inverse-transform sampling of the piecewise-constant pdf given by ``weights`` over ``bins``
    (run_nerf_helpers.py:264-307): bins [R,nb], weights [R,nb-1] -> samples [R,N_samples]."""
    lead = bins.shape[:-1]
    b2 = bins.reshape(-1, bins.shape[-1])
    w2 = weights.reshape(-1, weights.shape[-1])
    u, u_det = _uniform_variates(b2.shape[0], N_samples, det, pytest, b2.device)
    out = ops.sample_pdf(b2, w2, N_samples, u=u, u_det=u_det)
    return out.reshape(*lead, N_samples)


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """Volume-render raw network outputs (run_nerf_helpers.py:577-628).

    raw [R,S,4], z_vals [R,S], rays_d [R,3] ->
    (rgb_map [R,3], disp_map [R], acc_map [R], weights [R,S], depth_map [R], sparsity_loss [R])."""
    noise = None
    if raw_noise_std > 0.:
        if pytest:  # :603-606 (uniform variates, as the reference)
            np.random.seed(0)
            noise = torch.Tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std).to(raw.device)
        else:
            noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
    return ops.CompositeFn.apply(raw, z_vals, rays_d, noise, white_bkgd)


def render_rays(ray_batch,
                network_fn,
                network_query_fn,
                N_samples,
                embed_fn=None,
                retraw=False,
                lindisp=False,
                perturb=0.,
                N_importance=0,
                network_fine=None,
                white_bkgd=False,
                raw_noise_std=0.,
                verbose=False,
                pytest=False):
    """Coarse pass, importance resampling, fine pass (run_nerf_helpers.py:464-574).

    ray_batch [R, 8 or 11] = (origin, direction, near, far[, unit view direction]).  Returns the
    reference's dict: rgb_map, depth_map, acc_map, sparsity_loss[, raw][, rgb0, depth0, acc0,
    sparsity_loss0, z_std].

    With ``HN_AUTO_GRAPH=1`` (hn_b200.autograph) repeated training calls of one signature are replayed as two CUDA
    graphs (forward, backward) behind one autograd node; everything else runs the statements below."""
    if (autograph.ENABLED and torch.is_grad_enabled() and ray_batch.is_cuda and not pytest and not DEBUG
            and not verbose):
        kw = dict(network_fn=network_fn, network_query_fn=network_query_fn, N_samples=N_samples, embed_fn=embed_fn,
                  retraw=retraw, lindisp=lindisp, perturb=perturb, N_importance=N_importance,
                  network_fine=network_fine, white_bkgd=white_bkgd, raw_noise_std=raw_noise_std)
        ret = autograph.render_rays(_render_rays_eager, ray_batch, kw)
        if ret is not None:
            return ret
    return _render_rays_eager(ray_batch, network_fn, network_query_fn, N_samples, embed_fn, retraw, lindisp, perturb,
                              N_importance, network_fine, white_bkgd, raw_noise_std, verbose, pytest)


def _render_rays_eager(ray_batch, network_fn, network_query_fn, N_samples, embed_fn=None, retraw=False, lindisp=False,
                       perturb=0., N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.,
                       verbose=False, pytest=False):
    rb = ray_batch if (ray_batch.dtype == torch.float32 and ray_batch.is_contiguous()) else ray_batch.float().contiguous()
    R, width = rb.shape
    rays_o, rays_d = rb[:, 0:3], rb[:, 3:6]
    rays_d_c = rays_d.contiguous()                       # one packed copy shared by both compositing passes
    viewdirs = rb[:, -3:].contiguous() if width > 8 else None
    dev = rb.device

    t_vals = _linspace01(N_samples, dev)
    t_rand = None
    if perturb > 0.:
        if pytest:  # :531-534
            np.random.seed(0)
            t_rand = torch.Tensor(np.random.rand(R, N_samples)).to(dev)
        else:
            t_rand = torch.rand(R, N_samples, device=dev)
    z_vals = ops.coarse_z(rb[:, 6], rb[:, 7], width, t_vals, t_rand, R, N_samples, lindisp)   # :514-536
    pts = ops.ray_points(rays_o, rays_d, width, z_vals)                                       # :538

    raw = network_query_fn(pts, viewdirs, network_fn)
    rgb_map, disp_map, acc_map, weights, depth_map, sparsity_loss = raw2outputs(
        raw, z_vals, rays_d_c, raw_noise_std, white_bkgd, pytest=pytest)

    if N_importance > 0:
        rgb_map_0, depth_map_0, acc_map_0, sparsity_loss_0 = rgb_map, depth_map, acc_map, sparsity_loss
        # :547-552 and :568 in one launch: mids, sample_pdf(weights[..., 1:-1]), sort(cat), std
        det = (perturb == 0.)
        u, u_det = _uniform_variates(R, N_importance, det, pytest, dev)
        z_samples, z_vals, z_std = ops.resample(z_vals, weights.detach(), N_importance, u=u, u_det=u_det)
        pts = ops.ray_points(rays_o, rays_d, width, z_vals)                                   # :552
        run_fn = network_fn if network_fine is None else network_fine
        raw = network_query_fn(pts, viewdirs, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map, sparsity_loss = raw2outputs(
            raw, z_vals, rays_d_c, raw_noise_std, white_bkgd, pytest=pytest)

    ret = {'rgb_map': rgb_map, 'depth_map': depth_map, 'acc_map': acc_map, 'sparsity_loss': sparsity_loss}
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'] = rgb_map_0
        ret['depth0'] = depth_map_0
        ret['acc0'] = acc_map_0
        ret['sparsity_loss0'] = sparsity_loss_0
        ret['z_std'] = z_std
    if DEBUG:
        for k, v in ret.items():
            if torch.isnan(v).any() or torch.isinf(v).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


# ----------------------------------------------------------------------------------------------
# image-level drivers (run_nerf_helpers.py:310-459)
# ----------------------------------------------------------------------------------------------
def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1.,
           use_viewdirs=False, c2w_staticcam=None, **kwargs):
    """Render a full image (``c2w``) or a given ray batch (``rays`` = (rays_o, rays_d)).

    Returns [rgb_map, depth_map, acc_map, extras]."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays

    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)

    out_shape = rays_d.shape
    if ndc:
        rays_o, rays_d = get_ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3])
    rays_d = torch.reshape(rays_d, [-1, 3])
    if rays_d.is_cuda and isinstance(near, (int, float)) and isinstance(far, (int, float)):
        # normalise the view directions and build [o | d | near | far | viewdir] in one launch
        packed = ops.pack_rays(rays_o, rays_d, None if viewdirs is None else torch.reshape(viewdirs, [-1, 3]),
                               near, far)
    else:
        rays_o, rays_d = rays_o.float(), rays_d.float()
        ones = torch.ones_like(rays_d[..., :1])
        columns = [rays_o, rays_d, near * ones, far * ones]
        if use_viewdirs:
            viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
            columns.append(torch.reshape(viewdirs, [-1, 3]).float())
        packed = torch.cat(columns, -1)

    pieces = {}
    for i in range(0, packed.shape[0], chunk):
        for k, v in render_rays(packed[i:i + chunk], **kwargs).items():
            pieces.setdefault(k, []).append(v)
    merged = {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in pieces.items()}
    for k in merged:
        merged[k] = torch.reshape(merged[k], list(out_shape[:-1]) + list(merged[k].shape[1:]))

    head = ['rgb_map', 'depth_map', 'acc_map']
    return [merged[k] for k in head] + [{k: v for k, v in merged.items() if k not in head}]


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0):
    """Render every pose of a camera path; optionally score against ground truth and save figures."""
    H, W, focal = hwf
    near, far = render_kwargs['near'], render_kwargs['far']
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor

    rgbs, depths, psnrs = [], [], []
    t0 = time.time()
    for i, c2w in enumerate(render_poses):
        print(i, time.time() - t0)
        t0 = time.time()
        rgb, depth, acc, _ = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], **render_kwargs)
        rgbs.append(rgb.cpu().numpy())
        depths.append(((depth - near) / (far - near)).cpu().numpy())
        if i == 0:
            print(rgb.shape, depth.shape)
        if gt_imgs is not None and render_factor == 0:
            gt = gt_imgs[i]
            gt = gt.cpu().numpy() if isinstance(gt, torch.Tensor) else gt
            p = -10. * np.log10(np.mean(np.square(rgbs[-1] - gt)))
            print(p)
            psnrs.append(p)
        if savedir is not None:
            _save_figure(os.path.join(savedir, '{:03d}.png'.format(i)), to8b(rgbs[-1]), depths[-1])

    rgbs, depths = np.stack(rgbs, 0), np.stack(depths, 0)
    if gt_imgs is not None and render_factor == 0:
        avg_psnr = sum(psnrs) / len(psnrs)
        print("Avg PSNR over Test set: ", avg_psnr)
        if savedir is not None:
            with open(os.path.join(savedir, "test_psnrs_avg{:0.2f}.pkl".format(avg_psnr)), "wb") as fp:
                pickle.dump(psnrs, fp)
    return rgbs, depths


def _save_figure(filename, rgb8, depth01):
    """RGB next to the normalised depth map, as the reference's matplotlib figure (:436-447)."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except (ImportError, AttributeError):  # no (usable) matplotlib: fall back to a side-by-side PNG via PIL
        from PIL import Image
        d8 = to8b(np.repeat(depth01[..., None], 3, axis=-1))
        Image.fromarray(np.concatenate([rgb8, d8], axis=1)).save(filename)
        return
    fig = plt.figure(figsize=(25, 15))
    ax = fig.add_subplot(1, 2, 1)
    ax.imshow(rgb8)
    ax.axis('off')
    ax = fig.add_subplot(1, 2, 2)
    ax.imshow(depth01, cmap='plasma', vmin=0, vmax=1)
    ax.axis('off')
    plt.savefig(filename, bbox_inches='tight', pad_inches=0)
    plt.close(fig)
