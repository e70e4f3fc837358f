"""Ray generation -- drop-in for the functions of the reference's ``ray_util.py`` that ``run_nerf.py`` and
``render`` use (get_rays :62-80, get_rays_np :82-93, get_ndc_rays :96-142), plus the two helpers its ``bbox.py``
imports from this module (get_directions :8-33, ray_from_directions :35-60).

These feed the hot path but are not on it ("next" row 3).  ``get_rays`` on a CUDA pose and
``get_ndc_rays`` on CUDA rays are one kernel launch each (hn_get_rays, hn_ndc_rays); the numpy variant and CPU
tensors stay plain array math.  The kornia-based equirectangular helpers (ray_util.py:8-57) serve the st3d branch, which is
dead in the reference (Appendix B6), and are not provided.
"""
from __future__ import annotations

import numpy as np
import torch


def get_directions(H, W, focal):
    """Ray directions of all pixels in the camera frame, [H, W, 3] on the CPU (ray_util.py:8-33; used by the
    reference's bbox.py, which run_nerf.py's Blender / LLFF loaders call).  The reference builds the pixel grid
    with kornia.create_meshgrid, i.e. CPU linspaces whatever the default tensor type is."""
    xs = torch.linspace(0, W - 1, W, device="cpu", dtype=torch.float32)
    ys = torch.linspace(0, H - 1, H, device="cpu", dtype=torch.float32)
    j, i = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)


def ray_from_directions(directions, c2w):
    """World-space origins and NORMALISED directions of all pixels, each [H*W, 3] (ray_util.py:35-60)."""
    rays_d = directions @ c2w[:3, :3].T
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def get_rays(H, W, K, c2w):
    """Pinhole camera rays in world space: (rays_o, rays_d), each [H, W, 3]."""
    if isinstance(c2w, torch.Tensor) and c2w.is_cuda:
        from hn_b200 import ops
        rays_d = ops.get_rays_d(H, W, K[0][0], K[1][1], K[0][2], K[1][2], c2w)
        return c2w[:3, -1].expand(rays_d.shape), rays_d
    dev = c2w.device if isinstance(c2w, torch.Tensor) else None
    xs = torch.linspace(0, W - 1, W, device=dev)
    ys = torch.linspace(0, H - 1, H, device=dev)
    j, i = torch.meshgrid(ys, xs, indexing="ij")  # i: column (x), j: row (y), both [H, W]
    cam = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], dim=-1)
    rays_d = torch.sum(cam[..., None, :] * c2w[:3, :3], dim=-1)  # rotate into the world frame
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_rays_np(H, W, K, c2w):
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    cam = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], axis=-1)
    rays_d = np.sum(cam[..., np.newaxis, :] * c2w[:3, :3], axis=-1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def get_ndc_rays(H, W, focal, near, rays_o, rays_d):
    """Forward-facing scenes: move origins to the near plane and map to normalised device coordinates.
    CUDA rays with a scalar near plane: one launch (hn_ndc_rays), bit-identical to the chain below."""
    if isinstance(rays_d, torch.Tensor) and rays_d.is_cuda and isinstance(near, (int, float)) \
            and rays_o.shape == rays_d.shape:
        from hn_b200 import ops
        return ops.ndc_rays(H, W, focal, near, rays_o, rays_d)
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz = rays_o[..., 0] / rays_o[..., 2]
    oy_oz = rays_o[..., 1] / rays_o[..., 2]
    sx = -1.0 / (W / (2.0 * focal))
    sy = -1.0 / (H / (2.0 * focal))
    o = torch.stack([sx * ox_oz, sy * oy_oz, 1.0 + 2.0 * near / rays_o[..., 2]], dim=-1)
    d = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - ox_oz),
                     sy * (rays_d[..., 1] / rays_d[..., 2] - oy_oz),
                     1 - o[..., 2]], dim=-1)
    return o, d
