"""Opt-in CUDA graphs UNDER the drop-in API (``HN_AUTO_GRAPH=1`` or ``autograph.enable()``).

An unmodified ``run_nerf.py`` drives ``render`` -> ``render_rays`` eagerly: at N_rand = 1024 the step is bound by
the host (about 60 launches forward and 40 backward issued from Python and the autograd engine), not by the GPU.
``GraphedTrainStep`` removes that, but it needs the caller to hand over its loop.  This module keeps the caller's
loop: once a ``render_rays`` call with the same shapes / options / networks has been seen a few times with
gradients enabled, its forward is captured into one CUDA graph and the backward of its outputs into a second one
(the scheme of ``torch.cuda.make_graphed_callables``), and later calls replay them behind one autograd node:

    ret = render_rays(batch, ...)      # copy batch -> static buffer, replay the forward graph
    loss = img2mse(ret['rgb_map'], target) + ...      # the caller's own loss, eager
    loss.backward()                    # copy the output gradients -> static buffers, replay the backward graph

Parameter gradients are accumulated by the captured kernels straight into the modules' persistent GradSink
buffers, as in the eager path.  Outputs are STATIC buffers: a call's results are overwritten by the next call with
the same signature (a training loop consumes them within the iteration).  The backward graph is captured per SET of
outputs that actually received a gradient (first backward with a new set: the captured forward's autograd graph is
run as it is; the graph for that set is captured before the next forward): outputs the loss does not use must not
become roots with zero gradients -- an empty ray's disparity is NaN and 0 x NaN would poison its gradient.

Two things make this safe under a foreign loop:
* everything runs on ONE non-default stream that this module makes current (``ensure_stream``) and on which it
  also captures.  A leaf's AccumulateGrad node belongs to the stream it was created on and lives as long as any
  autograd graph references it -- and the caller still holds last iteration's ``loss`` when it calls ``render``
  again -- so a capture on a separate stream would have to synchronise with uncaptured work
  (cudaErrorStreamCaptureIsolation).  With one stream there is nothing to synchronise with.
* a forward replay while the previous one's backward is still pending (two forwards, then two backwards) would
  clobber the saved activations; such a call, and any call whose tensors moved (re-homed parameters, a new gradient
  buffer), runs eagerly / drops the captured graphs.

Anything unexpected during capture disables the entry and the eager path carries on.
"""
from __future__ import annotations

import os
import warnings
from typing import Callable, Dict, List, Optional

import torch

from . import ops

ENABLED = os.environ.get("HN_AUTO_GRAPH", "0") == "1"
WARMUP_CALLS = 3          # eager calls per signature before the capture (lazy allocations, optimizer plans ...)

_stream: Optional[torch.cuda.Stream] = None
_entries: Dict[tuple, "_Entry"] = {}
stats = {"captures": 0, "captures_bwd": 0, "replays": 0, "eager": 0, "failed": 0}


def enable(on: bool = True) -> None:
    global ENABLED
    ENABLED = bool(on)
    if not on:
        _entries.clear()


def ensure_stream(device) -> Optional[torch.cuda.Stream]:
    """Make one non-default stream current (once per process) and return it."""
    global _stream
    if not ENABLED or not torch.cuda.is_available():
        return None
    if _stream is None:
        dev = torch.device(device)
        _stream = torch.cuda.Stream(device=dev)
        _stream.wait_stream(torch.cuda.current_stream(dev))
        torch.cuda.set_stream(_stream)
    return _stream


def reset() -> None:
    """Drop every captured graph (tests; after changing model structure)."""
    _entries.clear()


def shutdown() -> None:
    """Disable, drop the graphs and make the default stream current again."""
    global _stream, ENABLED
    ENABLED = False
    _entries.clear()
    if _stream is not None:
        dev = _stream.device
        torch.cuda.default_stream(dev).wait_stream(_stream)
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        _stream = None


class _Entry:
    __slots__ = ("calls", "failed", "pending", "g_f", "bwd", "want_bwd", "static_in", "keys", "outs", "diff_idx",
                 "static_grads", "sinks", "ptrs", "params", "slots", "owners")

    def __init__(self):
        self.calls = 0
        self.failed = False
        self.pending = False
        self.g_f = None
        self.bwd = {}
        self.want_bwd = None


def _modules_of(kw) -> List[torch.nn.Module]:
    mods = []
    for name in ("network_fn", "network_fine", "embed_fn"):
        m = kw.get(name)
        if isinstance(m, torch.nn.Module) and all(m is not o for o in mods):
            mods.append(m)
    return mods


def _sinks_of(mods) -> list:
    sinks = []
    for m in mods:
        s = m.grad_sink() if hasattr(m, "grad_sink") else getattr(m, "_sink", None)
        if s is not None:
            sinks.append(s)
    return sinks


def _pointers(params, sinks) -> tuple:
    return tuple(p.data_ptr() for p in params) + tuple(s.flat.data_ptr() if s.flat is not None else 0 for s in sinks)


class _ReplayFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, entry, ray_batch, *params):
        ctx.set_materialize_grads(False)
        entry.static_in.copy_(ray_batch, non_blocking=True)
        entry.g_f.replay()
        entry.pending = True
        ctx.entry = entry
        outs = tuple(o.detach() for o in entry.outs)
        ctx.mark_non_differentiable(*[o for i, o in enumerate(outs) if i not in entry.diff_idx])
        return outs

    @staticmethod
    def backward(ctx, *grads):
        entry = ctx.entry
        # Only the outputs the caller's loss uses are roots of the backward pass.  (Feeding zeros for the others is
        # NOT the same: an empty ray has acc = 0, its disparity is 1 / (depth / acc) = NaN, and 0 x NaN poisons the
        # ray's gradient -- autograd never visits an output without a gradient, and neither does the eager path.)
        used = tuple(i for i in entry.diff_idx if grads[i] is not None)
        if not used:
            entry.pending = False
            return (None, None) + (None,) * len(entry.params)
        for i in used:
            entry.static_grads[i].copy_(grads[i], non_blocking=True)
        for s in entry.sinks:      # zero_grad(set_to_none=True) bookkeeping: re-attach param.grad, clear if needed
            s.acquire()
        g_b = entry.bwd.get(used)
        if g_b is not None:
            g_b.replay()
        else:
            # first backward with this set of outputs: run the captured forward's autograd graph as it is (its saved
            # tensors are the static buffers the forward replay has just filled); render_rays() captures the graph
            # for this set before the next forward
            torch.autograd.backward([entry.outs[i] for i in used], [entry.static_grads[i] for i in used],
                                    retain_graph=True)
            entry.want_bwd = used
        entry.pending = False
        return (None, None) + (None,) * len(entry.params)


def _capture_backward(entry: _Entry, used: tuple, stream) -> None:
    g_b = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_b, stream=stream, pool=entry.g_f.pool()):
        torch.autograd.backward([entry.outs[i] for i in used], [entry.static_grads[i] for i in used], retain_graph=True)
    entry.bwd[used] = g_b
    stats["captures_bwd"] = stats.get("captures_bwd", 0) + 1


def _capture(entry: _Entry, impl: Callable, ray_batch: torch.Tensor, kw: dict, mods, params) -> None:
    dev = ray_batch.device
    stream = torch.cuda.current_stream(dev)
    for hook in ops.pre_capture_hooks:
        hook()
    # the sinks must exist and own param.grad before the capture, so that the captured backward contains no zero-fill
    sinks = _sinks_of(mods)
    for s in sinks:
        s.acquire()
    entry.static_in = ray_batch.detach().clone()
    g_f = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_f, stream=stream):
        with torch.enable_grad():
            ret = impl(entry.static_in, **kw)
    keys = list(ret.keys())
    outs = [ret[k] for k in keys]
    diff_idx = [i for i, o in enumerate(outs) if o.requires_grad]
    if not diff_idx:
        raise RuntimeError("no differentiable output")
    sinks = _sinks_of(mods)    # the forward may have (re)created them
    entry.g_f = g_f
    entry.keys, entry.outs, entry.diff_idx = keys, outs, diff_idx
    entry.static_grads = {i: torch.zeros_like(outs[i]) for i in diff_idx}
    entry.bwd, entry.want_bwd = {}, None
    entry.sinks, entry.params = sinks, params
    entry.ptrs = _pointers(params, sinks)
    stats["captures"] += 1


def _param_slots(mods):
    """[(owning submodule, name)] of every trainable parameter: read back per call with two dict lookups each
    (nn.Module.parameters() walks the module tree: ~0.1 ms per call for the 40 parameters of a NeRF)."""
    slots = []
    for m in mods:
        for sub in m.modules():
            for name, p in sub._parameters.items():
                if p is not None and p.requires_grad:
                    slots.append((sub, name))
    return slots


def render_rays(impl: Callable, ray_batch: torch.Tensor, kw: dict) -> Optional[dict]:
    """The graphed ``render_rays`` for this call, or None when the caller should run ``impl`` eagerly."""
    if _stream is None or torch.cuda.current_stream(ray_batch.device) != _stream or ray_batch.requires_grad:
        return None
    mods = _modules_of(kw)
    key = (tuple(ray_batch.shape), ray_batch.dtype, id(kw.get("network_query_fn")), tuple(id(m) for m in mods),
           kw.get("N_samples"), kw.get("N_importance"), float(kw.get("perturb", 0.)), bool(kw.get("white_bkgd")),
           float(kw.get("raw_noise_std", 0.)), bool(kw.get("lindisp")), bool(kw.get("retraw")))
    entry = _entries.get(key)
    if entry is None:
        entry = _Entry()
        entry.slots = _param_slots(mods)
        if not entry.slots:
            return None            # nothing to train: not a training call
        # the key holds ids: keep the objects alive as long as the entry, so that an id cannot be recycled into it
        entry.owners = (tuple(mods), kw.get("network_query_fn"))
        _entries[key] = entry
    entry.calls += 1
    if entry.failed:
        return None
    params = [sub._parameters[name] for sub, name in entry.slots]
    if entry.g_f is None:
        if entry.calls <= WARMUP_CALLS:
            stats["eager"] += 1
            return None
        try:
            _capture(entry, impl, ray_batch, kw, mods, params)
        except Exception as exc:  # noqa: BLE001 -- whatever it was, the eager path still works
            entry.failed = True
            entry.g_f = None
            stats["failed"] += 1
            warnings.warn(f"hn_b200.autograph: capture failed ({type(exc).__name__}: {exc}); this call signature "
                          "stays eager")
            torch.cuda.synchronize(ray_batch.device)
            return None
    elif any(a is not b for a, b in zip(entry.params, params)) or entry.ptrs != _pointers(params, _sinks_of(mods)):
        del _entries[key]          # parameters replaced or tensors moved: capture again after the next warm-up
        stats["eager"] += 1
        return None
    if entry.want_bwd is not None and not entry.pending:
        used, entry.want_bwd = entry.want_bwd, None
        if used not in entry.bwd:
            try:
                for hook in ops.pre_capture_hooks:
                    hook()
                for s in entry.sinks:
                    s.acquire()    # param.grad attached: the captured backward contains no zero-fill
                _capture_backward(entry, used, torch.cuda.current_stream(ray_batch.device))
            except Exception as exc:  # noqa: BLE001
                del _entries[key]
                stats["failed"] += 1
                warnings.warn(f"hn_b200.autograph: backward capture failed ({type(exc).__name__}: {exc})")
                torch.cuda.synchronize(ray_batch.device)
                return None
    if entry.pending:              # the previous forward's backward has not run: do not clobber its activations
        entry.pending = False
        stats["eager"] += 1
        return None
    outs = _ReplayFn.apply(entry, ray_batch, *entry.params)
    stats["replays"] += 1
    return dict(zip(entry.keys, outs))
