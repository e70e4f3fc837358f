"""Tensor-level wrappers over the C ABI (include/hashnerf_b200.h).

PyTorch is used here for device memory, streams and autograd bookkeeping only: every function
below marshals ``data_ptr()``/shape/stride plus the current CUDA stream into one C call.  There
is no alternative implementation: CPU tensors are rejected with a RuntimeError.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import weakref

import torch

from . import _lib

MLP_PARAMS = 9344
MLP_SHAPES = ((64, 32), (16, 64), (64, 31), (64, 64), (3, 64))  # W0..W4, nn.Linear.weight layout


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _need_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "hashnerf_b200 runs on CUDA (sm_100a) only: got a tensor on "
                f"{t.device}; there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    if dev is None:
        raise RuntimeError("no tensor argument")
    return dev


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter avoids building a Stream
    object per call; it is what the launch-overhead-bound eager path spends its time on otherwise)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


# bumped by every parameter update that goes through this package's raw-pointer kernels (RAdam.step, graph replay):
# torch's own version counters do not see those writes, caches keyed on parameter values check this instead
param_epoch = [0]
# callables run right before a CUDA-graph capture starts (graph.GraphedTrainStep): modules that cache tensors carrying
# an autograd graph between calls drop them here (see GraphedTrainStep._capture)
pre_capture_hooks = []


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(device: torch.device):
    """Make ``device`` current for the duration of a C call (a shared no-op when it already is)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def _consecutive(tensors: Sequence[torch.Tensor]) -> bool:
    """True if the (contiguous, fp32) tensors sit back to back in ONE storage (adjacent addresses of
    separate allocations do not count: a view spanning them would leave its storage)."""
    addr = tensors[0].data_ptr()
    base = tensors[0].untyped_storage().data_ptr()
    for t in tensors:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.data_ptr() != addr \
                or t.untyped_storage().data_ptr() != base:
            return False
        addr += t.numel() * 4
    return True


def pack(tensors: Sequence[torch.Tensor], known_flat: bool = False) -> torch.Tensor:
    """One flat fp32 view over ``tensors`` -- zero-copy when they already are slices of one flat buffer
    (the layout HashEmbedder / NeRFSmall allocate), otherwise a packed copy.  ``known_flat``: the caller has
    just verified the layout."""
    if known_flat or _consecutive(tensors):
        n = sum(t.numel() for t in tensors)
        return torch.as_strided(tensors[0].detach(), (n,), (1,))
    return torch.cat([t.detach().reshape(-1).float() for t in tensors])


# ----------------------------------------------------------------------------------------------
# hash encoding
# ----------------------------------------------------------------------------------------------
def spatial_hash(coords: torch.Tensor, log2_hashmap_size: int) -> torch.Tensor:
    """embedding/hash_encoding.py:112-128 on CUDA: int64 [..., dim] -> int64 [...]."""
    dev = _need_cuda(coords)
    c = coords.to(torch.int64).contiguous()
    out = torch.empty(c.shape[:-1], dtype=torch.int64, device=dev)
    n = out.numel()
    with _on(dev):
        _lib.call("hn_spatial_hash", c.data_ptr(), n, int(c.shape[-1]), int(log2_hashmap_size), out.data_ptr(),
                  _stream())
    return out


def voxel_vertices(x, bbox6, resolutions, log2_hashmap_size):
    """Parity hook: (hashed [L,N,8] int64, vmin [L,N,3], vmax [L,N,3]) for every level."""
    dev = _need_cuda(x, bbox6, resolutions)
    x = _f32c(x)
    N, L = x.shape[0], resolutions.numel()
    hashed = torch.empty(L, N, 8, dtype=torch.int64, device=dev)
    vmin = torch.empty(L, N, 3, dtype=torch.float32, device=dev)
    vmax = torch.empty(L, N, 3, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_voxel_vertices", x.data_ptr(), bbox6.data_ptr(), resolutions.data_ptr(), N, L,
                  int(log2_hashmap_size), hashed.data_ptr(), vmin.data_ptr(), vmax.data_ptr(), _stream())
    return hashed, vmin, vmax


def hash_encode_forward(x, tables_flat, bbox6, resolutions, L, F, log2T, want_keep=True):
    dev = _need_cuda(x, tables_flat, bbox6, resolutions)
    x = _f32c(x)
    N = x.shape[0]
    out = torch.empty(N, L * F, dtype=torch.float32, device=dev)
    keep = torch.empty(N, dtype=torch.uint8, device=dev) if want_keep else None
    with _on(dev):
        _lib.call("hn_hash_encode_fwd", x.data_ptr(), tables_flat.data_ptr(), bbox6.data_ptr(),
                  resolutions.data_ptr(), N, L, F, log2T, out.data_ptr(), _ptr(keep), _stream())
    return out, keep


def hash_encode_backward(x, dy, bbox6, resolutions, L, F, log2T, dtables_flat, ordered=False):
    """Accumulates into ``dtables_flat`` ([L * 2^T * F] fp32).  ``ordered``: the points are spatially coherent
    in their given order (ray samples) -> warp-aggregated scatter."""
    dev = _need_cuda(x, dy, bbox6, resolutions, dtables_flat)
    x, dy = _f32c(x), _f32c(dy)
    with _on(dev):
        _lib.call("hn_hash_encode_bwd_ordered" if ordered else "hn_hash_encode_bwd", x.data_ptr(), dy.data_ptr(), bbox6.data_ptr(), resolutions.data_ptr(),
                  x.shape[0], L, F, log2T, dtables_flat.data_ptr(), _stream())


def hash_sort_points(x, bbox6, grid_res: int = 128) -> torch.Tensor:
    """Counting sort of the points by grid cell -> xs4 [N,4] = (x, y, z, bit-cast original row)."""
    dev = _need_cuda(x, bbox6)
    x = _f32c(x)
    N = x.shape[0]
    lib = _lib.load()
    nbytes = lib.hn_hash_sort_workspace_bytes(N, int(grid_res))
    if nbytes < 0:
        raise RuntimeError("hn_hash_sort_workspace_bytes rejected its arguments")
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=dev)
    xs4 = torch.empty(N, 4, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_hash_sort_points", x.data_ptr(), bbox6.data_ptr(), N, int(grid_res), ws.data_ptr(),
                  xs4.data_ptr(), _stream())
    return xs4


def hash_encode_forward_sorted(xs4, tables_flat, bbox6, resolutions, L, F, log2T, want_keep=True):
    dev = _need_cuda(xs4, tables_flat, bbox6, resolutions)
    N = xs4.shape[0]
    out = torch.empty(N, L * F, dtype=torch.float32, device=dev)
    keep = torch.empty(N, dtype=torch.uint8, device=dev) if want_keep else None
    with _on(dev):
        _lib.call("hn_hash_encode_fwd_sorted", xs4.data_ptr(), tables_flat.data_ptr(), bbox6.data_ptr(),
                  resolutions.data_ptr(), N, L, F, log2T, out.data_ptr(), _ptr(keep), _stream())
    return out, keep


def hash_encode_backward_sorted(xs4, dy, bbox6, resolutions, L, F, log2T, dtables_flat, levels=None):
    """``levels = (begin, end)`` restricts the scatter to that level range (gradient buckets, see dp.py)."""
    dev = _need_cuda(xs4, dy, bbox6, resolutions, dtables_flat)
    dy = _f32c(dy)
    with _on(dev):
        if levels is None:
            _lib.call("hn_hash_encode_bwd_sorted", xs4.data_ptr(), dy.data_ptr(), bbox6.data_ptr(),
                      resolutions.data_ptr(), xs4.shape[0], L, F, log2T, dtables_flat.data_ptr(), _stream())
        else:
            _lib.call("hn_hash_encode_bwd_sorted_levels", xs4.data_ptr(), dy.data_ptr(), bbox6.data_ptr(),
                      resolutions.data_ptr(), xs4.shape[0], L, F, log2T, dtables_flat.data_ptr(), int(levels[0]),
                      int(levels[1]), _stream())


# Points are re-ordered by grid cell before encoding when there are at least this many of them (the sort
# costs a few passes over the points; measured crossover for uniformly random points: 2^18 plain 0.27 ms vs
# sorted 0.31 ms, 2^19 0.52 vs 0.42 ms, 2^22 3.78 vs 2.20 ms -- tools/exp_threshold.py).
SORT_MIN_POINTS = 1 << 19


def sort_grid_res(n_points: int) -> int:
    """About one point per cell, rounded to a power of two so that cell faces coincide with the voxel faces of the
    power-of-two levels (measured, tools/exp_grid.py / exp_grid2.py: 2^24 points -> 256 is a sharp optimum of the
    scatter, 2^22 -> 128 beats the cube root 161, 2^20 -> 128 beats 102), capped at the two-level sort's limit."""
    import math
    exp = round(math.log2(max(n_points, 1)) / 3.0)
    return int(min(256, max(16, 1 << exp)))


class GradSink:
    """Persistent flat gradient buffer behind a group of parameters that are views of one flat buffer.

    ``torch.autograd`` would hand every backward call a fresh dense gradient per parameter and add them up
    (for the 16 level tables: a 64 MiB memset plus 16 adds per pass).  Instead the backward kernels accumulate
    straight into this buffer and ``param.grad`` is pointed at its slices, which is also what lets RAdam and
    the data-parallel all-reduce treat the group as ONE tensor.  Semantics are those of autograd: gradients
    accumulate until ``zero_grad()``; both ``set_to_none=True`` (buffer re-zeroed on the next backward) and
    ``set_to_none=False`` (slices zeroed in place) work, and gradients produced by other autograd paths
    (e.g. the TV loss through ``nn.Embedding``) are merged."""

    _by_ptr = {}  # flat.data_ptr() -> weakref(sink): lets the optimizer report a buffer it cleared in its own pass

    def __init__(self, params):
        self.params = list(params)
        self.flat = None
        self._views = None  # one view of `flat` per parameter, created once (param.grad is set to these objects)
        self.clean = False  # the whole buffer is known to be zero (cleared by a fused optimizer pass)

    @classmethod
    def note_cleared(cls, ptr: int, numel: int) -> None:
        """Called by RAdam(fused_zero_grad=True) after a pass that zeroed ``numel`` gradients starting at ``ptr``."""
        ref = cls._by_ptr.get(ptr)
        sink = ref() if ref is not None else None
        if sink is not None and sink.flat is not None and sink.flat.data_ptr() == ptr and sink.flat.numel() == numel:
            sink.clean = True

    def _allocate(self):
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        GradSink._by_ptr[self.flat.data_ptr()] = weakref.ref(self)
        self._views, off = [], 0
        for p in self.params:
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def adopt(self, flat: torch.Tensor) -> None:
        """Use ``flat`` (zero-filled, one float per parameter element, e.g. a slice of a symmetric-memory buffer that
        peers can read) as the gradient buffer from now on."""
        if flat.numel() != sum(p.numel() for p in self.params) or flat.dtype != torch.float32:
            raise RuntimeError("adopted gradient buffer has the wrong size or dtype")
        self.flat = flat
        GradSink._by_ptr[flat.data_ptr()] = weakref.ref(self)
        self._views, off = [], 0
        for p in self.params:
            self._views.append(flat[off:off + p.numel()].view_as(p))
            p.grad = self._views[-1]
            off += p.numel()
        self.clean = True

    def acquire(self) -> torch.Tensor:
        """Flat fp32 buffer to accumulate into; afterwards every param.grad is a slice of it."""
        fresh = False
        if self.flat is None or self.flat.device != self.params[0].device:
            self._allocate()
            fresh = True
        views = self._views
        clean, self.clean = self.clean, False     # whatever happens next, gradients are about to be accumulated
        mine = [p.grad is v for p, v in zip(self.params, views)]
        if all(mine):
            return self.flat                      # still accumulating into our buffer (the common case)
        for i, (p, v) in enumerate(zip(self.params, views)):   # same memory through another tensor object
            if not mine[i] and p.grad is not None and p.grad.data_ptr() == v.data_ptr() and p.grad.shape == v.shape:
                mine[i] = True
        foreign = [(v, p.grad) for p, v, m in zip(self.params, views, mine) if (not m) and p.grad is not None]
        if not fresh:
            if any(mine):                         # mixed state: keep what is ours, clear the rest
                for v, m in zip(views, mine):
                    if not m:
                        v.zero_()
            elif not clean:
                self.flat.zero_()                 # first backward since zero_grad(set_to_none=True)
        for v, g in foreign:                      # gradients that arrived through plain autograd
            v.add_(g)
        for p, v in zip(self.params, views):
            p.grad = v
        return self.flat


class HashEncodeFn(torch.autograd.Function):
    """features, keep = HashEncodeFn.apply(x, bbox6, resolutions, log2T, F, coherent, sink, *level_tables)

    ``level_tables`` are the L ``nn.Embedding.weight`` parameters ([2^T, F] each).  Autograd routes the
    table gradient to each of them; the gradients returned are slices of ONE flat buffer filled by a
    single scatter kernel (the reference produces 16 separate dense gradients through
    embedding_dense_backward, hash_encoding.py:106).  ``coherent``: None = sort the points by grid cell when
    there are many of them, True / False = force, "ordered" = the points are already spatially coherent in the
    given order (consecutive samples of rays): never sort, aggregate the scatter.  ``sink``: a GradSink (gradients accumulate in place into its
    persistent buffer and ``param.grad`` points at it) or None (plain autograd return values)."""

    @staticmethod
    def forward(ctx, x, bbox6, resolutions, log2T, F, coherent, sink, *level_tables):
        L = len(level_tables)
        ctx.sink = sink
        flat = pack(level_tables)
        N = x.shape[0]
        ordered = isinstance(coherent, str) and coherent == "ordered"
        if ordered:
            use_sort = False
        else:
            use_sort = (N >= SORT_MIN_POINTS) if coherent is None else (bool(coherent) and N > 0)
        if use_sort:
            xs4 = hash_sort_points(x, bbox6, sort_grid_res(N))
            out, keep = hash_encode_forward_sorted(xs4, flat, bbox6, resolutions, L, F, log2T)
            ctx.save_for_backward(xs4, bbox6, resolutions)
        else:
            out, keep = hash_encode_forward(x, flat, bbox6, resolutions, L, F, log2T)
            ctx.save_for_backward(x, bbox6, resolutions)
        ctx.meta = (L, F, log2T, use_sort, ordered)
        ctx.mark_non_differentiable(keep)
        return out, keep  # keep: uint8 (HashEmbedder.forward converts to bool for the caller)

    @staticmethod
    def backward(ctx, dout, _dkeep):
        x, bbox6, resolutions = ctx.saved_tensors
        L, F, log2T, use_sort, ordered = ctx.meta
        if ctx.needs_input_grad[0]:
            # the reference is differentiable w.r.t. the points (through the interpolation weights); nothing on the
            # training path asks for that gradient and no kernel computes it -- say so instead of returning None
            raise RuntimeError("hashnerf_b200: the gradient of the hash encoding w.r.t. the input points is not "
                               "implemented (detach the points, or use the reference's encoder for that term)")
        T = 1 << log2T
        sink = ctx.sink
        dflat = sink.acquire() if sink is not None else torch.zeros(L * T * F, dtype=torch.float32, device=x.device)
        if use_sort:
            hash_encode_backward_sorted(x, dout, bbox6, resolutions, L, F, log2T, dflat)
        else:
            hash_encode_backward(x, dout, bbox6, resolutions, L, F, log2T, dflat, ordered=ordered)
        if sink is not None:  # accumulated in place; param.grad already points into the buffer
            return (None,) * (7 + L)
        grads = dflat.view(L, T, F).unbind(0)
        return (None, None, None, None, None, None, None) + tuple(grads)


class TVLossFn(torch.autograd.Function):
    """tv = TVLossFn.apply(weight [2^T, F], origin int64[3] (device), cube, log2T, sink_info|None)

    ``sink_info`` = (GradSink, element offset of this level in its flat buffer): the backward kernel then
    scatters straight into the persistent gradient buffer; otherwise a dense gradient is returned."""

    @staticmethod
    def forward(ctx, weight, origin, cube, log2T, sink_info):
        dev = _need_cuda(weight, origin)
        w = weight.detach()
        if not w.is_contiguous() or w.dtype != torch.float32:
            raise RuntimeError("TV loss expects a contiguous fp32 table")
        out = torch.empty((), dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.call("hn_tv_loss_fwd", w.data_ptr(), origin.data_ptr(), int(cube), int(log2T), int(w.shape[1]),
                      out.data_ptr(), _stream())
        ctx.save_for_backward(w, origin)
        ctx.meta = (int(cube), int(log2T), sink_info)
        return out

    @staticmethod
    def backward(ctx, gout):
        w, origin = ctx.saved_tensors
        cube, log2T, sink_info = ctx.meta
        gout = _f32c(gout)
        if sink_info is not None:
            sink, offset = sink_info
            target = sink.acquire()[offset:offset + w.numel()]
            ret = None
        else:
            target = torch.zeros_like(w)
            ret = target
        with _on(w.device):
            _lib.call("hn_tv_loss_bwd", w.data_ptr(), origin.data_ptr(), cube, log2T, int(w.shape[1]),
                      gout.data_ptr(), target.data_ptr(), _stream())
        return ret, None, None, None, None


class TVSweepFn(torch.autograd.Function):
    """tv[L] = TVSweepFn.apply(flat_tables, origins int64 [L,3], cubes int32 [L], max_cube, log2T, F, sink, *levels)

    The total-variation terms of ALL levels in one launch (forward) and one launch (backward).  ``flat_tables`` is
    the detached flat view over the level tables; ``levels`` are the level Parameters themselves (the autograd
    inputs).  With a GradSink the backward scatters straight into the persistent gradient buffer, otherwise it
    returns one dense gradient per level (slices of one buffer)."""

    @staticmethod
    def forward(ctx, flat_tables, origins, cubes, max_cube, log2T, F, sink, *levels):
        dev = _need_cuda(flat_tables, origins, cubes)
        L = len(levels)
        out = torch.empty(L, dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.call("hn_tv_loss_fwd_levels", flat_tables.data_ptr(), origins.data_ptr(), cubes.data_ptr(), L,
                      int(max_cube), int(log2T), int(F), out.data_ptr(), _stream())
        ctx.save_for_backward(flat_tables, origins, cubes)
        ctx.meta = (L, int(max_cube), int(log2T), int(F), sink)
        return out

    @staticmethod
    def backward(ctx, gout):
        flat_tables, origins, cubes = ctx.saved_tensors
        L, max_cube, log2T, F, sink = ctx.meta
        gout = _f32c(gout)
        if sink is not None:
            target = sink.acquire()
        else:
            target = torch.zeros(flat_tables.numel(), dtype=torch.float32, device=flat_tables.device)
        with _on(flat_tables.device):
            _lib.call("hn_tv_loss_bwd_levels", flat_tables.data_ptr(), origins.data_ptr(), cubes.data_ptr(), L,
                      max_cube, log2T, F, gout.data_ptr(), target.data_ptr(), _stream())
        if sink is not None:
            return (None,) * (7 + L)
        return (None,) * 7 + tuple(target.view(L, -1, F).unbind(0))


class TVSweepPartsFn(torch.autograd.Function):
    """The same launch as TVSweepFn, but the L terms come back as L separate scalars: a training loop that adds
    them up one by one (``sum(total_variation_loss(...) for i in range(16))``, run_nerf.py:628-635) then has no
    ``select`` node per level, whose backward is a zero-fill + copy + accumulate -- 48 tiny launches per step."""

    @staticmethod
    def forward(ctx, flat_tables, origins, cubes, max_cube, log2T, F, sink, *levels):
        ctx.set_materialize_grads(False)
        out = TVSweepFn.forward(ctx, flat_tables, origins, cubes, max_cube, log2T, F, sink, *levels)
        return tuple(out.unbind(0))

    @staticmethod
    def backward(ctx, *gouts):
        L = ctx.meta[0]
        given = [g for g in gouts if g is not None]
        if not given:
            return (None,) * (7 + L)
        if len(given) == L:
            gout = torch.stack([g.reshape(()) for g in gouts])
        else:
            zero = torch.zeros((), dtype=given[0].dtype, device=given[0].device)
            gout = torch.stack([zero if g is None else g.reshape(()) for g in gouts])
        return TVSweepFn.backward(ctx, gout)


# ----------------------------------------------------------------------------------------------
# spherical harmonics
# ----------------------------------------------------------------------------------------------
def sh_encode(dirs: torch.Tensor, degree: int) -> torch.Tensor:
    dev = _need_cuda(dirs)
    d = _f32c(dirs).reshape(-1, 3)
    out = torch.empty(d.shape[0], degree * degree, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_sh_encode", d.data_ptr(), d.shape[0], int(degree), out.data_ptr(), _stream())
    return out.reshape(*dirs.shape[:-1], degree * degree)


# ----------------------------------------------------------------------------------------------
# NeRFSmall
# ----------------------------------------------------------------------------------------------
def _rows(t: torch.Tensor, width: int) -> Tuple[torch.Tensor, int]:
    """A 2-D fp32 view whose rows are contiguous (last stride 1); returns (tensor, row stride)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2 or t.shape[1] != width:
        raise RuntimeError(f"expected a [N,{width}] tensor, got {tuple(t.shape)}")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < width):
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else width)


class MLPFn(torch.autograd.Function):
    """out[N,4] = MLPFn.apply(enc[N,32], views[Nv,16], pts_per_view, keep|None, sink|None, W0, W1, W2, W3, W4)"""

    @staticmethod
    def forward(ctx, enc, views, pts_per_view, keep, sink, *weights):
        ctx.sink = sink
        dev = _need_cuda(enc, views, *weights)
        enc_r, enc_stride = _rows(enc, 32)
        views_r, views_stride = _rows(views, 16)
        N = enc_r.shape[0]
        if views_r.shape[0] * pts_per_view < N:
            raise RuntimeError("views has too few rows for pts_per_view")
        wflat = pack(weights)
        keep_u8 = None if keep is None else (keep if keep.dtype == torch.uint8 else keep.to(torch.uint8)).contiguous()
        out = torch.empty(N, 4, dtype=torch.float32, device=dev)
        # what autograd would keep of the forward pass: the ReLU gates (24 bytes per point), only when a backward
        # pass can follow
        gates = torch.empty(N, 6, dtype=torch.int32, device=dev) if any(ctx.needs_input_grad) else None
        with _on(dev):
            _lib.call("hn_mlp_fwd", enc_r.data_ptr(), enc_stride, views_r.data_ptr(), views_stride,
                      int(pts_per_view), wflat.data_ptr(), _ptr(keep_u8), N, out.data_ptr(), _ptr(gates), _stream())
        ctx.gates = gates
        ctx.save_for_backward(enc_r, views_r, wflat, keep_u8 if keep_u8 is not None else torch.empty(0, device=dev))
        ctx.meta = (enc_stride, views_stride, int(pts_per_view), keep_u8 is not None, N,
                    ctx.needs_input_grad[0])
        return out

    @staticmethod
    def backward(ctx, dout):
        enc_r, views_r, wflat, keep_u8 = ctx.saved_tensors
        enc_stride, views_stride, ppv, has_keep, N, _ = ctx.meta
        dev = enc_r.device
        dout = _f32c(dout)
        d_enc = torch.empty(N, 32, dtype=torch.float32, device=dev)
        sink = ctx.sink
        dflat = sink.acquire() if sink is not None else torch.zeros(MLP_PARAMS, dtype=torch.float32, device=dev)
        lib = _lib.load()
        ws = torch.empty(max(1, lib.hn_mlp_bwd_workspace_bytes(N) // 4), dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.call("hn_mlp_bwd", enc_r.data_ptr(), enc_stride, views_r.data_ptr(), views_stride, ppv,
                      wflat.data_ptr(), keep_u8.data_ptr() if has_keep else None, _ptr(ctx.gates), dout.data_ptr(), N,
                      d_enc.data_ptr(), dflat.data_ptr(), ws.data_ptr(), _stream())
        if sink is not None:
            return (d_enc, None, None, None, None) + (None,) * 5
        grads, off = [], 0
        for (o, i) in MLP_SHAPES:
            grads.append(dflat[off:off + o * i].view(o, i))
            off += o * i
        return (d_enc, None, None, None, None) + tuple(grads)


# ----------------------------------------------------------------------------------------------
# photometric loss
# ----------------------------------------------------------------------------------------------
class MSEFn(torch.autograd.Function):
    """mean((a - b) ** 2) of two fp32 CUDA tensors of one shape: one launch forward, one backward."""

    @staticmethod
    def forward(ctx, a, b):
        dev = _need_cuda(a, b)
        a, b = _f32c(a), _f32c(b)
        out = torch.empty((), dtype=torch.float32, device=dev)
        with _on(dev):
            _lib.call("hn_mse_fwd", a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), _stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        need_a, need_b = ctx.needs_input_grad
        gout = _f32c(gout)
        da = torch.empty_like(a) if need_a else None
        db = torch.empty_like(b) if need_b else None
        with _on(a.device):
            _lib.call("hn_mse_bwd", a.data_ptr(), b.data_ptr(), a.numel(), gout.data_ptr(), _ptr(da), _ptr(db), _stream())
        return da, db


def mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return MSEFn.apply(a, b)


# ----------------------------------------------------------------------------------------------
# compositing
# ----------------------------------------------------------------------------------------------
class CompositeFn(torch.autograd.Function):
    """rgb, disp, acc, weights, depth, entropy = CompositeFn.apply(raw, z, rays_d, noise|None, white)"""

    @staticmethod
    def forward(ctx, raw, z, rays_d, noise, white_bkgd):
        # outputs the loss does not use (disp, acc, depth, ...) reach backward as None, not as zero tensors that a
        # fill kernel each had to produce; hn_composite_bwd takes NULL for them
        ctx.set_materialize_grads(False)
        dev = _need_cuda(raw, z, rays_d, noise)
        raw, z, rays_d = _f32c(raw), _f32c(z), _f32c(rays_d)
        noise = None if noise is None else _f32c(noise)
        R, S = z.shape
        if raw.shape != (R, S, 4):
            raise RuntimeError(f"raw must be [R,S,4], got {tuple(raw.shape)} for z {tuple(z.shape)}")
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        rgb, disp, acc, weights, depth, ent = new(R, 3), new(R), new(R), new(R, S), new(R), new(R)
        with _on(dev):
            _lib.call("hn_composite_fwd", raw.data_ptr(), z.data_ptr(), rays_d.data_ptr(), _ptr(noise), R, S,
                      int(bool(white_bkgd)), rgb.data_ptr(), disp.data_ptr(), acc.data_ptr(), weights.data_ptr(),
                      depth.data_ptr(), ent.data_ptr(), _stream())
        ctx.save_for_backward(raw, z, rays_d, noise if noise is not None else torch.empty(0, device=dev))
        ctx.meta = (R, S, bool(white_bkgd), noise is not None)
        return rgb, disp, acc, weights, depth, ent

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_weights, d_depth, d_ent):
        raw, z, rays_d, noise = ctx.saved_tensors
        R, S, white, has_noise = ctx.meta
        g = [None if t is None else _f32c(t) for t in (d_rgb, d_disp, d_acc, d_weights, d_depth, d_ent)]
        if all(t is None for t in g):
            return None, None, None, None, None
        d_raw = torch.empty_like(raw)
        with _on(raw.device):
            _lib.call("hn_composite_bwd", raw.data_ptr(), z.data_ptr(), rays_d.data_ptr(),
                      noise.data_ptr() if has_noise else None, R, S, int(white), _ptr(g[0]), _ptr(g[1]), _ptr(g[2]),
                      _ptr(g[3]), _ptr(g[4]), _ptr(g[5]), d_raw.data_ptr(), _stream())
        return d_raw, None, None, None, None


# ----------------------------------------------------------------------------------------------
# sampling (no gradients flow through these on the reference path: z_samples is detached, :549)
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def sample_pdf(bins, weights, n_samples: int, u: Optional[torch.Tensor] = None,
               u_det: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _need_cuda(bins, weights, u, u_det)
    bins, weights = _f32c(bins), _f32c(weights)
    R, nb = bins.shape
    if weights.shape != (R, nb - 1):
        raise RuntimeError(f"weights must be [R, nbins-1] = {(R, nb - 1)}, got {tuple(weights.shape)}")
    if u is not None:
        u = _f32c(u)
        if u.shape != (R, n_samples):
            raise RuntimeError("u must be [R, n_samples]")
    else:
        u_det = _f32c(u_det)
    out = torch.empty(R, n_samples, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_sample_pdf", bins.data_ptr(), weights.data_ptr(), _ptr(u), _ptr(u_det), R, nb, int(n_samples),
                  out.data_ptr(), _stream())
    return out


@torch.no_grad()
def resample(z, weights, n_importance: int, u: Optional[torch.Tensor] = None, u_det: Optional[torch.Tensor] = None):
    """(z_samples [R,Ni], z_merged [R,S+Ni], z_std [R]) -- the resampling block of render_rays in one launch."""
    dev = _need_cuda(z, weights, u, u_det)
    z, weights = _f32c(z), _f32c(weights)
    R, S = z.shape
    if u is not None:
        u = _f32c(u)
    else:
        u_det = _f32c(u_det)
    samples = torch.empty(R, n_importance, dtype=torch.float32, device=dev)
    merged = torch.empty(R, S + n_importance, dtype=torch.float32, device=dev)
    z_std = torch.empty(R, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_resample", z.data_ptr(), weights.data_ptr(), _ptr(u), _ptr(u_det), R, S, int(n_importance),
                  samples.data_ptr(), merged.data_ptr(), z_std.data_ptr(), _stream())
    return samples, merged, z_std


@torch.no_grad()
def sort_concat_rows(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    R = a.shape[0]
    out = torch.empty(R, a.shape[1] + b.shape[1], dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_sort_concat_rows", a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], R, out.data_ptr(),
                  _stream())
    return out


@torch.no_grad()
def coarse_z(near, far, nf_stride: int, t_vals, t_rand, R: int, S: int, lindisp: bool) -> torch.Tensor:
    dev = _need_cuda(near, far, t_vals, t_rand)
    z = torch.empty(R, S, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_coarse_z", near.data_ptr(), far.data_ptr(), int(nf_stride), t_vals.data_ptr(), _ptr(t_rand),
                  R, S, int(bool(lindisp)), z.data_ptr(), _stream())
    return z


@torch.no_grad()
def ray_points(rays_o, rays_d, ray_stride: int, z) -> torch.Tensor:
    dev = _need_cuda(rays_o, rays_d, z)
    R, S = z.shape
    pts = torch.empty(R, S, 3, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_ray_points", rays_o.data_ptr(), rays_d.data_ptr(), int(ray_stride), z.data_ptr(), R, S,
                  pts.data_ptr(), _stream())
    return pts


# ----------------------------------------------------------------------------------------------
# ray generation / packing (no gradients: cameras are data)
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def get_rays_d(H: int, W: int, fx: float, fy: float, cx: float, cy: float, c2w: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(c2w)
    m = c2w if (c2w.dtype == torch.float32 and c2w.stride(-1) == 1) else c2w.float().contiguous()
    out = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_get_rays", int(H), int(W), float(fx), float(fy), float(cx), float(cy), m.data_ptr(),
                  int(m.stride(0)), out.data_ptr(), _stream())
    return out


def ndc_rays(H: int, W: int, focal: float, near: float, rays_o: torch.Tensor, rays_d: torch.Tensor):
    """ray_util.py:96-142 in one launch: (rays_o, rays_d) [..., 3] -> NDC (o, d), same leading shape."""
    dev = _need_cuda(rays_o, rays_d)
    shape = rays_d.shape
    o, o_stride = _rows3(rays_o.reshape(-1, 3))
    d, d_stride = _rows3(rays_d.reshape(-1, 3))
    R = d.shape[0]
    out_o = torch.empty(R, 3, dtype=torch.float32, device=dev)
    out_d = torch.empty(R, 3, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_ndc_rays", int(H), int(W), float(focal), float(near), o.data_ptr(), o_stride, d.data_ptr(),
                  d_stride, R, out_o.data_ptr(), out_d.data_ptr(), _stream())
    return out_o.reshape(shape), out_d.reshape(shape)


def sample_rays(images: torch.Tensor, poses: torch.Tensor, H: int, W: int, K, near: float, far: float,
                step_params: torch.Tensor, n_rand: int, rays: Optional[torch.Tensor] = None,
                target: Optional[torch.Tensor] = None, want_pix: bool = False):
    """hn_sample_rays: images [n_img,H,W,C>=3] fp32 and poses [n_img,3,4] fp32 resident on the device;
    step_params int32[6] on the device = (image index, seed, row0, col0, win_h, win_w).  Returns
    (rays [n_rand,11], target [n_rand,3][, pix [n_rand,2] int32])."""
    dev = _need_cuda(images, poses, step_params)
    if images.dtype != torch.float32 or not images.is_contiguous() or images.dim() != 4 or images.shape[1:3] != (H, W):
        raise RuntimeError("images must be a contiguous fp32 [n_img, H, W, C] tensor")
    if poses.dtype != torch.float32 or not poses.is_contiguous() or poses.shape[1:] != (3, 4):
        raise RuntimeError("poses must be a contiguous fp32 [n_img, 3, 4] tensor")
    if step_params.dtype != torch.int32 or step_params.numel() < 6:
        raise RuntimeError("step_params must be int32[6] on the device")
    rays = torch.empty(n_rand, 11, dtype=torch.float32, device=dev) if rays is None else rays
    target = torch.empty(n_rand, 3, dtype=torch.float32, device=dev) if target is None else target
    pix = torch.empty(n_rand, 2, dtype=torch.int32, device=dev) if want_pix else None
    with _on(dev):
        _lib.call("hn_sample_rays", images.data_ptr(), int(images.shape[3]), poses.data_ptr(), int(H), int(W),
                  float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]), float(near), float(far),
                  step_params.data_ptr(), int(n_rand), rays.data_ptr(), target.data_ptr(), _ptr(pix), _stream())
    return (rays, target, pix) if want_pix else (rays, target)


def _rows3(t: torch.Tensor):
    if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
        t = t.float().contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else 3)


@torch.no_grad()
def pack_rays(rays_o, rays_d, viewdirs, near: float, far: float) -> torch.Tensor:
    dev = _need_cuda(rays_o, rays_d, viewdirs)
    o, so = _rows3(rays_o)
    d, sd = _rows3(rays_d)
    v, sv = (None, 0) if viewdirs is None else _rows3(viewdirs)
    R = d.shape[0]
    out = torch.empty(R, 11 if v is not None else 8, dtype=torch.float32, device=dev)
    with _on(dev):
        _lib.call("hn_pack_rays", o.data_ptr(), so, d.data_ptr(), sd, _ptr(v), sv, float(near), float(far), R,
                  out.data_ptr(), _stream())
    return out
