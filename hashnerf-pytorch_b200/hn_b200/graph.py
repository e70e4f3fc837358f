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
                 grad_sync: Optional[Callable[[], None]] = None, batcher=None, fused_zero_grad: bool = True):
        """render_fn(ray_batch[n_rays, ray_width]) -> dict (e.g. a closure over render_rays);
        loss_fn(ret, target[n_rays,3]) -> scalar; optimizer: radam.RAdam.  ``warmup`` eager steps are real
        optimisation steps (they also initialise the optimizer state and every lazy allocation).
        ``grad_sync``: called between backward and the optimizer, eagerly and inside the capture -- data-parallel
        runs pass ``dp.GradSync(...).all_reduce_inline`` (NCCL collectives on the capturing stream become graph
        nodes) and set ``optimizer.grad_scale = 1 / world_size``.
        ``batcher``: a ``batcher.DeviceRayBatcher``; ``step()`` then takes no arguments and the pixel sampling / ray
        generation / target gather launch is part of the captured graph.
        ``fused_zero_grad``: the optimizer pass also clears the gradients it consumed (radam.RAdam.fused_zero_grad),
        so neither the eager warm-up steps nor the graph contain a separate 64 MiB fill of the table gradient."""
        self._grad_sync = grad_sync
        self._batcher = batcher
        self.opt = optimizer
        if fused_zero_grad and hasattr(optimizer, "fused_zero_grad"):
            optimizer.fused_zero_grad = True
        self.lr_schedule = lr_schedule
        self.global_step = 0
        self.rays = torch.zeros(n_rays, ray_width, device=device)
        self.target = torch.zeros(n_rays, 3, device=device)
        if batcher is not None:    # the batcher's launch writes straight into the graph's static buffers
            if batcher.n_rand != n_rays:
                raise RuntimeError("batcher.n_rand must equal n_rays")
            batcher.rays, batcher.target = self.rays, self.target
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

    def _fill(self):
        """This step's rays and targets into the static buffers (batcher launch: also what the graph replays)."""
        rays, target = self._batcher.launch()
        if rays.data_ptr() != self.rays.data_ptr():
            self.rays.copy_(rays, non_blocking=True)
            self.target.copy_(target, non_blocking=True)

    def step(self, rays: Optional[torch.Tensor] = None, target: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on (rays, target) -- or on the batcher's next draw; returns the (device) loss."""
        if self._batcher is not None:
            self._batcher.prepare()
        else:
            self.rays.copy_(rays, non_blocking=True)
            self.target.copy_(target, non_blocking=True)
        self._set_lr()
        self.global_step += 1
        if self.graph is None:
            if self._batcher is not None:
                self._fill()
            if self._warmup > 0:
                self._warmup -= 1
                self.loss = self._eager().detach()
                return self.loss
            # the eager step inside _capture() IS this call's optimisation step; from the next call on self.loss is
            # the graph's static output, refreshed by every replay
            return self._capture()
        self.opt.graph_prepare()
        self.graph.replay()
        ops.param_epoch[0] += 1
        return self.loss

    def _capture(self):
        """One last eager step (on a side stream, so that every lazy allocation and the optimizer's span plan
        exist), then the capture.  Capturing executes nothing: this batch is applied exactly once and the
        optimizer's step counters stay in lockstep with ``global_step``."""
        side = torch.cuda.Stream(device=self.rays.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager_loss = self._eager().detach()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # No autograd graph over the parameters may be alive when the capture starts: a leaf's AccumulateGrad node
        # belongs to the stream it was created on and lives as long as any graph references it, so a graph kept from
        # an eager step (e.g. the TV sweep loss.py caches between calls) would make the captured backward synchronise
        # with that uncaptured stream -- cudaErrorStreamCaptureIsolation.  Modules that cache such tensors register a
        # hook that drops them; callers must not hold on to losses of earlier steps either (detach them).
        for hook in ops.pre_capture_hooks:
            hook()
        self.opt.graph_plan()
        self.opt.zero_grad(set_to_none=True)   # the captured backward must start with "no gradient yet"
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if self._batcher is not None:
                self._fill()
            loss = self._loss(self._render(self.rays), self.target)
            loss.backward()
            if self._grad_sync is not None:
                self._grad_sync()
            self.opt.graph_launch()
            self.loss = loss.detach()
        return eager_loss
