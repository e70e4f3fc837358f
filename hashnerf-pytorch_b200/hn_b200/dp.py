"""Data parallelism over rays.

The reference has no distributed code at all (train.sh launches independent runs).  Rays are independent
units and the parameters are small enough to replicate (64 MiB of tables + 2 x 37 KB of MLP weights at
T = 2^19), so the only exchange step is one summed all-reduce of the gradients per training step:

* every rank renders its own ray batch (``shard_range`` for a shared batch, or its own sampled rays);
* ``GradSync.all_reduce()`` sums the gradients over ranks -- the 16 level tables are one flat buffer and the
  five matrices of each NeRFSmall another, so this is 3 NCCL calls, not 26 -- on a side stream;
* ``GradSync.wait()`` joins the side stream, and the 1/world_size averaging is folded into the optimizer
  kernel (``RAdam.grad_scale``) instead of a separate pass over 64 MiB.

Plumbing only: ``torch.distributed`` (NCCL on GPUs; the unit tests drive the same code over gloo on CPU
tensors with world_size 2).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of ``n_items`` rays for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _flat_runs(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Group tensors that sit back to back in one storage into single flat views (zero copy)."""
    runs, cur = [], []
    for t in tensors:
        if cur and ops._consecutive([cur[-1], t]):
            cur.append(t)
        else:
            if cur:
                runs.append(cur)
            cur = [t]
    if cur:
        runs.append(cur)
    out = []
    for run in runs:
        if len(run) == 1:
            out.append(run[0])
        else:
            n = sum(t.numel() for t in run)
            out.append(torch.as_strided(run[0], (n,), (1,)))
    return out


def broadcast_parameters(params: Iterable[torch.Tensor], src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters (identical init)."""
    with torch.no_grad():
        for flat in _flat_runs([p.data for p in params]):
            dist.broadcast(flat, src=src, group=group)


class GradSync:
    """Summed all-reduce of ``params``' gradients, flat-buffer aware.

    Usage per step::

        loss.backward()
        sync.all_reduce()          # enqueued on a side stream (CUDA) right after backward
        ...                        # anything that does not touch gradients overlaps here
        sync.wait()
        optimizer.grad_scale = sync.grad_scale   # 1 / world_size, applied inside the fused RAdam kernel
        optimizer.step()
    """

    def __init__(self, params: Iterable[torch.Tensor], group=None):
        self.params = [p for p in params]
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.grad_scale = 1.0 / self.world_size
        self._stream: Optional[torch.cuda.Stream] = None
        self._works = []
        self.bytes_last = 0
        self.calls_last = 0

    def _grads(self) -> List[torch.Tensor]:
        return [p.grad for p in self.params if p.grad is not None]

    def all_reduce(self) -> None:
        if self.world_size == 1:
            return
        grads = self._grads()
        if not grads:
            return
        flats = _flat_runs(grads)
        self.bytes_last = sum(f.numel() * f.element_size() for f in flats)
        self.calls_last = len(flats)
        if grads[0].is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=grads[0].device)
            self._stream.wait_stream(torch.cuda.current_stream(grads[0].device))
            with torch.cuda.stream(self._stream):
                for f in flats:
                    dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group)
                    f.record_stream(self._stream)
        else:  # gloo / CPU tensors (unit tests of the host logic)
            self._works = [dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for f in flats]

    def wait(self) -> None:
        if self._stream is not None:
            torch.cuda.current_stream(self._stream.device).wait_stream(self._stream)
        for w in self._works:
            w.wait()
        self._works = []

    def all_reduce_inline(self) -> None:
        """The same summed all-reduce issued on the CURRENT stream (no side stream, nothing to wait for): the form a
        CUDA-graph capture needs (graph.GraphedTrainStep(grad_sync=...))."""
        if self.world_size == 1:
            return
        flats = _flat_runs(self._grads())
        self.bytes_last = sum(f.numel() * f.element_size() for f in flats)
        self.calls_last = len(flats)
        for f in flats:
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group)


class BucketedTableReducer:
    """All-reduce of the hash-table gradient in level buckets, overlapped with the scatter.

    The flat gradient ``[L * 2^T * F]`` is level-major, so the levels ``[b, e)`` are the contiguous slice
    ``flat[b * slab : e * slab]``.  The caller scatters bucket after bucket on its compute stream
    (``ops.hash_encode_backward_sorted(..., levels=(b, e))``) and calls :meth:`reduce_levels` after each: the summed
    all-reduce of that slice is enqueued on a side stream behind an event, so it runs while the next bucket is
    being scattered.  :meth:`wait` joins the side stream.  On CPU tensors (gloo, the unit tests) the collectives
    are asynchronous works joined by :meth:`wait`.
    """

    def __init__(self, n_levels: int, group=None):
        self.n_levels = int(n_levels)
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.grad_scale = 1.0 / self.world_size
        self._stream: Optional[torch.cuda.Stream] = None
        self._works = []
        self.calls_last = 0

    @staticmethod
    def buckets(n_levels: int, levels_per_bucket: int = 4) -> List[Tuple[int, int]]:
        """[(begin, end)] covering [0, n_levels) in runs of ``levels_per_bucket`` (the last may be shorter)."""
        step = max(1, int(levels_per_bucket))
        return [(b, min(b + step, n_levels)) for b in range(0, n_levels, step)]

    def reduce_levels(self, flat: torch.Tensor, begin: int, end: int) -> None:
        if not (0 <= begin <= end <= self.n_levels):
            raise ValueError(f"level range [{begin}, {end}) outside [0, {self.n_levels})")
        if flat.numel() % self.n_levels:
            raise ValueError("flat gradient length is not a multiple of the level count")
        if self.world_size == 1 or begin == end:
            return
        slab = flat.numel() // self.n_levels
        piece = flat.view(-1)[begin * slab:end * slab]
        self.calls_last += 1
        if piece.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=piece.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(piece.device))   # the bucket's scatter has been enqueued
            self._stream.wait_event(ev)
            with torch.cuda.stream(self._stream):
                dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self._works.append(dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self) -> None:
        if self._stream is not None:
            torch.cuda.current_stream(self._stream.device).wait_stream(self._stream)
        for w in self._works:
            w.wait()
        self._works = []
        self.calls_last = 0


def all_gather_rows(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Inference: concatenate per-rank row blocks (``counts[r]`` rows from rank r) on every rank."""
    world = dist.get_world_size(group)
    width = local.shape[1:]
    pad = max(counts)
    buf = local.new_zeros((pad,) + tuple(width))
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


@torch.no_grad()
def render_image_sharded(H, W, K, c2w, render_fn, group=None):
    """Inference partition of SURVEY 8e: every rank renders a contiguous range of the frame's H*W rays and the
    per-ray results are all-gathered, so each rank returns the full (rgb [H,W,3], depth [H,W], acc [H,W]).

    ``render_fn(rays_o [n,3], rays_d [n,3]) -> (rgb [n,3], depth [n], acc [n])`` is typically a closure over
    ``run_nerf_helpers.render(H, W, K, rays=(o, d), ...)``."""
    from ray_util import get_rays
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rays_o, rays_d = get_rays(H, W, K, c2w)
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    start, stop = shard_range(H * W, rank, world)
    rgb, depth, acc = render_fn(rays_o[start:stop], rays_d[start:stop])
    local = torch.cat([rgb.reshape(-1, 3), depth.reshape(-1, 1), acc.reshape(-1, 1)], dim=-1)
    if world > 1:
        counts = [shard_range(H * W, r, world)[1] - shard_range(H * W, r, world)[0] for r in range(world)]
        local = all_gather_rows(local.contiguous(), counts, group=group)
    return local[:, :3].reshape(H, W, 3), local[:, 3].reshape(H, W), local[:, 4].reshape(H, W)
