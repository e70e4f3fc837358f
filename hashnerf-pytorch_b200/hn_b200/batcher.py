"""On-device training ray batcher (SURVEY section 8f "next" row 3) -- opt-in, next to the drop-in API.

The reference's Blender branch (run_nerf.py:576-605) does, per iteration: pick an image on the host, upload it
(1.9 MB at 400x400), generate the rays of ALL its pixels (160 k rays to pick 1024), build the pixel-coordinate grid,
draw N_rand distinct indices with numpy, upload them, gather rays and targets, and later (run_nerf_helpers.py:344-366)
normalise the view directions and pack the ray batch: ~25 launches, two uploads and host RNG on the step's critical
path.  ``DeviceRayBatcher`` keeps the images and poses resident on the device and does all of it in ONE launch
(hn_sample_rays) whose per-step scalars (image index, permutation seed, crop window) are read from device memory --
so the launch can sit inside a captured CUDA graph (``graph.GraphedTrainStep(..., batcher=...)``).

Pixels are drawn without replacement, as ``np.random.choice(..., replace=False)`` does at :600, through a keyed
permutation of the window's pixels; the stream of random numbers is of course not numpy's (the reference never
seeds torch and seeds numpy with 0, :30 -- no result depends on a particular stream).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class DeviceRayBatcher:
    _RING = 4

    def __init__(self, images, poses, H: int, W: int, K, near: float, far: float, n_rand: int, device,
                 i_train=None, precrop_iters: int = 0, precrop_frac: float = 0.5, seed: int = 0):
        """images [n_img,H,W,3|4] (numpy or tensor, already composited as run_nerf.py:262-266 leaves them),
        poses [n_img,>=3,4]; i_train: indices to draw images from (default: all)."""
        self.H, self.W, self.K, self.near, self.far, self.n_rand = int(H), int(W), K, float(near), float(far), int(n_rand)
        self.device = torch.device(device)
        img = torch.as_tensor(np.asarray(images) if not isinstance(images, torch.Tensor) else images)
        self.images = img.to(self.device, torch.float32).contiguous()
        pos = torch.as_tensor(np.asarray(poses) if not isinstance(poses, torch.Tensor) else poses)
        self.poses = pos.to(self.device, torch.float32)[:, :3, :4].contiguous()
        self.i_train = np.arange(self.images.shape[0]) if i_train is None else np.asarray(i_train)
        self.precrop_iters, self.precrop_frac = int(precrop_iters), float(precrop_frac)
        self.rng = np.random.RandomState(seed)
        self.rays = torch.empty(self.n_rand, 11, device=self.device)
        self.target = torch.empty(self.n_rand, 3, device=self.device)
        self._dev = torch.zeros(8, dtype=torch.int32, device=self.device)
        self._host = [torch.zeros(8, dtype=torch.int32).pin_memory() for _ in range(self._RING)]
        self._events = [None] * self._RING
        self._slot = 0
        self.step_index = 0

    def window(self, step: int):
        """(row0, col0, win_h, win_w): the centre crop of run_nerf.py:586-596 while step < precrop_iters."""
        H, W = self.H, self.W
        if step < self.precrop_iters:
            dH, dW = int(H // 2 * self.precrop_frac), int(W // 2 * self.precrop_frac)
            return H // 2 - dH, W // 2 - dW, 2 * dH, 2 * dW
        return 0, 0, H, W

    def prepare(self):
        """Host side of one step: draw the image and the permutation seed, hand them to the device through a guarded
        ring of pinned rows (the host may run ahead of the GPU; a row is rewritten only after its copy has run)."""
        slot = self._slot
        self._slot = (slot + 1) % self._RING
        if self._events[slot] is not None:
            self._events[slot].synchronize()
        row0, col0, wh, ww = self.window(self.step_index)
        if wh * ww < self.n_rand:
            raise RuntimeError(f"cannot draw {self.n_rand} distinct pixels from a {wh}x{ww} window")
        h = self._host[slot]
        h[0] = int(self.rng.choice(self.i_train))
        h[1] = int(self.rng.randint(0, 2 ** 31 - 1))
        h[2], h[3], h[4], h[5] = row0, col0, wh, ww
        with torch.cuda.device(self.device):
            self._dev.copy_(h, non_blocking=True)
            ev = self._events[slot] or torch.cuda.Event()
            ev.record()
            self._events[slot] = ev
        self.step_index += 1

    def launch(self, want_pix: bool = False):
        """The kernel only (graph-capturable): fills and returns (rays [n_rand,11], target [n_rand,3])."""
        return ops.sample_rays(self.images, self.poses, self.H, self.W, self.K, self.near, self.far, self._dev,
                               self.n_rand, rays=self.rays, target=self.target, want_pix=want_pix)

    def next(self, want_pix: bool = False):
        self.prepare()
        return self.launch(want_pix)
