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

``FusedExchange`` is the other form of the same step: parameters and gradients live in symmetric memory and one
kernel per flat buffer (csrc/dp_exchange.cu) reduces the gradients through the NVSwitch (multimem.ld_reduce) or over
peer NVLink loads, applies RAdam to the rank's 1/world slice with sharded moments, multicasts the new parameters and
clears the gradients -- collective, optimizer and zero_grad in a single pass.
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


def slice_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the slice of an n-float flat buffer that ``rank`` owns in hn_dp_reduce_update: chunks of
    ceil(n / world) rounded up to 4 floats (16-byte vectors), the last ranks possibly empty.  Mirrors
    csrc/dp_exchange.cu."""
    chunk = ((n + world_size - 1) // world_size + 3) & ~3
    begin = min(rank * chunk, n)
    return begin, min(begin + chunk, n)


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


class SymmetricBuffer:
    """A flat fp32 buffer allocated identically on every rank with peer (and, on NVSwitch systems, multicast)
    mappings.  torch.distributed._symmetric_memory does the allocation and the handle exchange; this package only
    takes the raw pointers from it."""

    def __init__(self, numel: int, device, group=None, dtype=torch.float32):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.tensor = symm.empty(int(numel), dtype=dtype, device=device)
        self.tensor.zero_()
        try:
            self.handle = symm.rendezvous(self.tensor, group=group)
        except Exception:  # older torch wants the group announced first
            symm.enable_symm_mem_for_group(group.group_name)
            self.handle = symm.rendezvous(self.tensor, group=group)
        self.rank, self.world = int(self.handle.rank), int(self.handle.world_size)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)   # 0: no NVSwitch multicast here
        self.itemsize = self.tensor.element_size()

    def ptr_table(self, offset_elems: int = 0) -> torch.Tensor:
        """int64 device tensor of the `world` peer pointers to element ``offset_elems`` of the buffer."""
        off = int(offset_elems) * self.itemsize
        return torch.tensor([p + off for p in self.peer_ptrs], dtype=torch.int64, device=self.tensor.device)

    def mc_ptr(self, offset_elems: int = 0) -> int:
        return (self.multicast_ptr + int(offset_elems) * self.itemsize) if self.multicast_ptr else 0


class FusedExchange:
    """Gradient exchange + optimizer + zero_grad as one pass over peer memory (csrc/dp_exchange.cu).

        fx = FusedExchange(optimizer, [emb, coarse, fine])      # once, after building the model on every rank
        ...
        loss.backward()          # gradients accumulate in the symmetric buffer (the modules' GradSinks point there)
        fx.step()                # replaces  sync.all_reduce(); sync.wait(); optimizer.step(); optimizer.zero_grad()

    ``modules``: a HashEmbedder and NeRFSmall networks (anything with ``_level_weights()`` / ``_weights()`` and a
    flat gradient sink).  Their parameters are re-homed, in order, into ONE symmetric parameter buffer and their
    gradient sinks into ONE symmetric gradient buffer; hyper-parameters (lr, betas, eps, weight_decay) are read from
    the optimizer's param groups every step, so the lr schedule run_nerf.py applies to them keeps working, but
    ``optimizer.step()`` / ``zero_grad()`` are not called any more.  The moments are SHARDED: each rank keeps
    exp_avg / exp_avg_sq only for the slice of every buffer it owns (``gather_moments()`` collects them for a
    checkpoint).  Parameters come out bit-identical on every rank: each slice is computed once, by its owner."""

    _RING = 4

    def __init__(self, optimizer, modules, group=None):
        if not dist.is_initialized():
            raise RuntimeError("FusedExchange needs an initialised torch.distributed process group (NCCL)")
        self.opt = optimizer
        self.group = group
        spans = []
        for mod in modules:
            ws = mod._level_weights() if hasattr(mod, "_level_weights") else mod._weights()
            if hasattr(mod, "_flatten_parameters"):
                mod._flatten_parameters()
            if not ops._consecutive(ws):
                raise RuntimeError("module parameters are not one flat buffer")
            spans.append((mod, ws, sum(w.numel() for w in ws)))
        dev = spans[0][1][0].device
        offs, total = [], 0
        for _m, _ws, n in spans:
            if n % 4:
                raise RuntimeError("flat parameter buffers must be multiples of 4 floats")
            offs.append(total)
            total += n
        self.params_sym = SymmetricBuffer(total, dev, group)
        self.grads_sym = SymmetricBuffer(total, dev, group)
        self.signals = SymmetricBuffer(256, dev, group, dtype=torch.int32)
        self.rank, self.world = self.params_sym.rank, self.params_sym.world
        self.multicast = bool(self.params_sym.multicast_ptr and self.grads_sym.multicast_ptr)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._spans, self._sinks = [], []
        with torch.no_grad():
            for (mod, ws, n), off in zip(spans, offs):
                flat_p = self.params_sym.tensor[off:off + n]
                flat_g = self.grads_sym.tensor[off:off + n]
                o = 0
                for w in ws:                                   # re-home the parameters (Parameter identity kept)
                    view = flat_p[o:o + w.numel()].view_as(w)
                    view.copy_(w)
                    w.data = view
                    o += w.numel()
                sink = mod.grad_sink() if hasattr(mod, "grad_sink") else None
                if sink is None:
                    if getattr(mod, "_sink", None) is None or any(a is not b for a, b in zip(mod._sink.params, ws)):
                        mod._sink = ops.GradSink(ws)
                    sink = mod._sink
                sink.adopt(flat_g)
                self._sinks.append(sink)
                gi = next(i for i, g in enumerate(optimizer.param_groups) if any(p is ws[0] for p in g['params']))
                self._spans.append(dict(off=off, n=n, group=gi, step=0,
                                        g_tab=self.grads_sym.ptr_table(off), p_tab=self.params_sym.ptr_table(off),
                                        g_mc=self.grads_sym.mc_ptr(off), p_mc=self.params_sym.mc_ptr(off)))
        self._sig_tab = self.signals.ptr_table(0)
        self._hp_dev = torch.zeros(len(self._spans), 8, dtype=torch.float32, device=dev)
        self._hp_host = [torch.zeros(len(self._spans), 8, dtype=torch.float32).pin_memory() for _ in range(self._RING)]
        self._events = [None] * self._RING
        self._slot = 0
        self._epoch = 0
        dist.barrier(group=group)          # every rank's buffers are initialised before anybody exchanges
        torch.cuda.synchronize(dev)

    def _prepare(self):
        slot = self._slot
        self._slot = (slot + 1) % self._RING
        if self._events[slot] is not None:
            self._events[slot].synchronize()
        host = self._hp_host[slot]
        for i, sp in enumerate(self._spans):
            grp = self.opt.param_groups[sp['group']]
            sp['step'] += 1
            beta1, beta2 = grp['betas']
            mode, step_size = self.opt._rectification(sp['step'], beta1, beta2)
            row = host[i]
            row[0], row[1], row[2] = beta1, beta2, grp['eps']
            row[3] = grp['weight_decay'] * grp['lr']
            row[4] = step_size * grp['lr']
            row[5], row[6], row[7] = 1.0 / self.world, float(mode), 0.0
        self._hp_dev.copy_(host, non_blocking=True)
        ev = self._events[slot] or torch.cuda.Event()
        ev.record()
        self._events[slot] = ev

    def _barrier(self, slot: int):
        _lib_call("hn_dp_barrier", self._sig_tab.data_ptr(), self.rank, self.world, slot, self._epoch, ops._stream())

    def step(self):
        """barrier -> (reduce + RAdam + broadcast + clear) per flat buffer -> barrier, all on the current stream."""
        dev = self.m.device
        with ops._on(dev):
            self._prepare()
            self._epoch += 1
            self._barrier(0)
            for i, sp in enumerate(self._spans):
                _lib_call("hn_dp_reduce_update", sp['g_tab'].data_ptr(), sp['g_mc'] or None, sp['p_tab'].data_ptr(),
                          sp['p_mc'] or None, self.m[sp['off']:].data_ptr(), self.v[sp['off']:].data_ptr(), self.rank,
                          self.world, sp['n'], self._hp_dev[i].data_ptr(), ops._stream())
            self._barrier(1)
        for sink in self._sinks:
            sink.clean = True              # the kernel cleared every rank's gradient buffer
        ops.param_epoch[0] += 1

    @torch.no_grad()
    def gather_moments(self):
        """(exp_avg, exp_avg_sq) of the whole parameter buffer on every rank (for checkpoints): each rank contributes
        the slices it owns."""
        m, v = self.m.clone(), self.v.clone()
        for sp in self._spans:
            n, off = sp['n'], sp['off']
            for t in (m, v):
                for r in range(self.world):
                    b, e = slice_bounds(n, r, self.world)
                    if e > b:
                        dist.broadcast(t[off + b:off + e], src=r, group=self.group)
        return m, v


class SymmetricAllReduce:
    """The summed all-reduce of one flat symmetric buffer through hn_dp_reduce_update(mode = -1): every rank reduces
    its slice through the switch and multicasts the sum (bench.py's N > 1 step; NCCL stays the reference point)."""

    def __init__(self, numel: int, device, group=None):
        self.buf = SymmetricBuffer(numel, device, group)
        self.signals = SymmetricBuffer(256, device, group, dtype=torch.int32)
        self.rank, self.world = self.buf.rank, self.buf.world
        self._g_tab, self._sig_tab = self.buf.ptr_table(0), self.signals.ptr_table(0)
        self._hp = torch.tensor([0, 0, 0, 0, 0, 1.0, -1.0, 0], dtype=torch.float32, device=device)
        self._epoch = 0
        self._tabs = {0: self._g_tab}   # peer-pointer tables per range start
        self.multicast = bool(self.buf.multicast_ptr)
        dist.barrier(group=group)
        torch.cuda.synchronize(device)

    @property
    def tensor(self) -> torch.Tensor:
        return self.buf.tensor

    def all_reduce(self, begin: int = 0, end: Optional[int] = None):
        """Sum elements [begin, end) of the buffer over the ranks, on the current stream.  Every rank must make the
        same sequence of calls (the barrier epochs count calls); begin and end - begin are multiples of 4."""
        n = self.buf.tensor.numel()
        end = n if end is None else int(end)
        begin = int(begin)
        if not (0 <= begin <= end <= n) or (begin & 3) or ((end - begin) & 3):
            raise ValueError(f"range [{begin}, {end}) must lie in [0, {n}) with begin and length multiples of 4")
        if end == begin:
            return
        tab = self._tabs.get(begin)
        if tab is None:
            tab = self._tabs[begin] = self.buf.ptr_table(begin)
        self._epoch += 1
        s = ops._stream()
        _lib_call("hn_dp_barrier", self._sig_tab.data_ptr(), self.rank, self.world, 0, self._epoch, s)
        _lib_call("hn_dp_reduce_update", tab.data_ptr(), self.buf.mc_ptr(begin) or None, None, None, None, None,
                  self.rank, self.world, end - begin, self._hp.data_ptr(), s)
        _lib_call("hn_dp_barrier", self._sig_tab.data_ptr(), self.rank, self.world, 1, self._epoch, s)


class OverlappedTableReducer:
    """The table-gradient exchange in level buckets over symmetric memory, overlapped with the scatter.

    The flat gradient ``[L * 2^T * F]`` is level-major; after the scatter of levels ``[b, e)`` has been enqueued
    (``ops.hash_encode_backward_sorted(..., levels=(b, e))``) :meth:`reduce_levels` enqueues the one-pass exchange
    of that slice (hn_dp_barrier + hn_dp_reduce_update) on a high-priority side stream behind an event, so it runs
    through the switch while the compute stream scatters the next bucket; only the last bucket's exchange is
    exposed.  All exchanges go through ONE side stream in call order, which keeps the barrier epochs in step on
    every rank.  :meth:`wait` joins the side stream."""

    def __init__(self, sar: SymmetricAllReduce, n_levels: int):
        self.sar, self.n_levels = sar, int(n_levels)
        dev = sar.tensor.device
        self._stream = torch.cuda.Stream(device=dev, priority=-1)   # above the compute stream: its few CTAs go first
        self._ev = torch.cuda.Event()

    def reduce_levels(self, begin: int, end: int) -> None:
        if not (0 <= begin <= end <= self.n_levels):
            raise ValueError(f"level range [{begin}, {end}) outside [0, {self.n_levels})")
        slab = self.sar.tensor.numel() // self.n_levels
        self._ev.record(torch.cuda.current_stream(self._stream.device))   # the bucket's scatter has been enqueued
        self._stream.wait_event(self._ev)
        with torch.cuda.stream(self._stream):
            self.sar.all_reduce(begin * slab, end * slab)

    def wait(self) -> None:
        torch.cuda.current_stream(self._stream.device).wait_stream(self._stream)


def _lib_call(name, *args):
    from . import _lib
    _lib.call(name, *args)


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
