"""Whole-training-step CUDA graph.

One NeRF training step at N_rand = 1024 is ~100 kernel launches of a few microseconds each: issued from
Python it is bound by launch overhead, not by the GPU.  ``GraphedTrainStep`` captures

    render_rays (coarse + fine)  ->  loss  ->  backward  ->  RAdam

once into a CUDA graph and replays it per step; rays and targets are copied into static buffers, the
step-dependent optimizer scalars travel through a pinned-host -> device copy node (radam.RAdam.graph_*).
Random draws (stratified jitter, resampling variates, sigma noise) use torch's graph-safe Philox generator,
so every replay draws fresh numbers.  This is an opt-in API next to the drop-in one: ``run_nerf.py`` drives
the same kernels eagerly.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, n_rays: int, render_fn: Callable[[torch.Tensor], dict], loss_fn: Callable[[dict, torch.Tensor], torch.Tensor],
                 optimizer, device, ray_width: int = 11, warmup: int = 3, lr_schedule: Optional[Callable[[int], float]] = None,
                 grad_sync: Optional[Callable[[], None]] = None):
        """render_fn(ray_batch[n_rays, ray_width]) -> dict (e.g. a closure over render_rays);
        loss_fn(ret, target[n_rays,3]) -> scalar; optimizer: radam.RAdam.  ``warmup`` eager steps are real
        optimisation steps (they also initialise the optimizer state and every lazy allocation).
        ``grad_sync``: called between backward and the optimizer, eagerly and inside the capture -- data-parallel
        runs pass ``dp.GradSync(...).all_reduce_inline`` (NCCL collectives on the capturing stream become graph
        nodes) and set ``optimizer.grad_scale = 1 / world_size``."""
        self._grad_sync = grad_sync
        self.opt = optimizer
        self.lr_schedule = lr_schedule
        self.global_step = 0
        self.rays = torch.zeros(n_rays, ray_width, device=device)
        self.target = torch.zeros(n_rays, 3, device=device)
        self._render, self._loss = render_fn, loss_fn
        self._warmup = warmup
        self.graph = None
        self.loss = None

    def _eager(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self._loss(self._render(self.rays), self.target)
        loss.backward()
        if self._grad_sync is not None:
            self._grad_sync()
        self.opt.step()
        return loss

    def _set_lr(self):
        if self.lr_schedule is not None:
            lr = self.lr_schedule(self.global_step)
            for g in self.opt.param_groups:
                g['lr'] = lr

    def step(self, rays: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """One optimisation step on (rays, target); returns the (device) loss of this step."""
        self.rays.copy_(rays, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self._set_lr()
        self.global_step += 1
        if self.graph is None and self._warmup > 0:
            self._warmup -= 1
            self.loss = self._eager().detach()
            return self.loss
        if self.graph is None:
            self._capture()
        self.opt.graph_prepare()
        self.graph.replay()
        ops.param_epoch[0] += 1
        return self.loss

    def _capture(self):
        side = torch.cuda.Stream(device=self.rays.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):      # one more eager step on the side stream so gradients exist for the plan
            self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.graph_plan()
        self.opt.zero_grad(set_to_none=True)   # the captured backward must start with "no gradient yet"
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss = self._loss(self._render(self.rays), self.target)
            loss.backward()
            if self._grad_sync is not None:
                self._grad_sync()
            self.opt.graph_launch()
            self.loss = loss.detach()
