"""Regularisers -- drop-in for the reference's ``loss.py`` (total_variation_loss :11-43,
sigma_sparsity_loss :45-47).  Adjacent to the hot path ("next" row 2): one CUDA launch for
the forward (hash, gather, squared differences, reduction) and one for the backward, instead of ~15 ATen
launches per level.
"""
from __future__ import annotations

import weakref
from math import exp, floor, log

import torch

from embedding.hash_encoding import hash, level_owner  # noqa: F401  (hash: reference import line loss.py:8)
from hn_b200 import ops


def _level_cube(min_resolution, max_resolution, level, n_levels):
    """(grid resolution, cube size) of one level -- fp64 ``math`` arithmetic exactly as loss.py:13-22.

    NOTE the resolution here is fp64 while the encoder's is fp32 tensor arithmetic; they can disagree (SURVEY
    Appendix B10) and both are kept as the reference has them."""
    growth = exp((log(max_resolution) - log(min_resolution)) / (n_levels - 1))
    resolution = int(floor(min_resolution * growth ** level))
    smallest = int(min_resolution) - 1
    largest = 50
    if smallest > largest:
        raise ValueError("ALERT! min cuboid size greater than max!")  # the reference drops into pdb here
    cube = int(floor(min(max(resolution / 10.0, smallest), largest)))
    return resolution, cube


# The reference's training loop asks for the 16 levels one after the other, every step (run_nerf.py:628-635):
# 16 random draws, 16 forward and 16 backward launches and, in Python, 16 autograd Functions -- more host time
# than the rest of a 1024-ray step.  When the level modules belong to a HashEmbedder of this package, the first
# request of such a sweep computes ALL levels (one draw, one launch each way, one autograd node) and the following
# requests, as long as they ask for increasing levels of the same unchanged tables, are served from that result.
# Set to False to evaluate every request on its own (one launch per level, the previous behaviour).
TV_SWEEP = True
# False (default): a sweep draws its 16 cube origins with the reference's own 16 ``torch.randint`` calls, in order, so
# a seeded run consumes the generator exactly as the reference does.  True: one ``torch.rand(L, 3)`` scaled to the
# same integer ranges (same distribution, another stream): 3 launches instead of 17 -- for CUDA-graph steps, where
# every node counts.
TV_FAST_DRAWS = False


class _Sweep:
    __slots__ = ("key", "vec", "last", "versions", "ptrs", "epoch")


_SWEEPS = weakref.WeakKeyDictionary()    # HashEmbedder -> its current sweep (kept off the module: no pickling issues)
_GEOMETRY = weakref.WeakKeyDictionary()  # HashEmbedder -> cached per-level cube sizes / spans on the device
ops.pre_capture_hooks.append(_SWEEPS.clear)   # a cached sweep keeps an autograd graph alive across a graph capture


def _sweep_terms(owner, key, min_resolution, max_resolution, log2_hashmap_size, n_levels, parts=False):
    """All levels' TV terms of ``owner`` as one [L] tensor -- or, ``parts``, as L scalars behind one autograd node --
    (a fresh random cube per level, loss.py:25)."""
    geo = _GEOMETRY.get(owner)
    flat = owner.flat_tables()
    dev = flat.device
    if geo is None or geo[0] != (key[:4], dev):
        res_cube = [_level_cube(min_resolution, max_resolution, l, n_levels) for l in range(n_levels)]
        spans = [r - c for r, c in res_cube]
        if min(spans) <= 0:
            raise RuntimeError("total_variation_loss: cube does not fit the level grid (random_(0, to<=0))")
        geo = ((key[:4], dev), spans,
               torch.tensor([c for _, c in res_cube], dtype=torch.int32, device=dev),
               max(c for _, c in res_cube), torch.tensor(spans, dtype=torch.float32, device=dev)[:, None])
        _GEOMETRY[owner] = geo
    _, spans, cubes, max_cube, span_t = geo
    if TV_FAST_DRAWS:
        origins = torch.minimum(torch.rand(len(spans), 3, device=dev) * span_t, span_t - 1).to(torch.int64)
    else:
        # the same draws, in the same order, as the reference's 16 consecutive calls make (loss.py:25): the generator
        # stream -- and with it a seeded training run -- stays identical to the reference's
        origins = torch.stack([torch.randint(0, sp, (3,), device=dev) for sp in spans])
    levels = owner._level_weights()
    sink = owner.grad_sink() if (torch.is_grad_enabled() and levels[0].requires_grad) else None
    fn = ops.TVSweepPartsFn if parts else ops.TVSweepFn
    return fn.apply(flat, origins, cubes, max_cube, int(log2_hashmap_size), int(flat.shape[-1]), sink, *levels)


def total_variation_sweep(embed_fn):
    """Opt-in: the TV terms of ALL levels of a HashEmbedder as one [L] tensor behind ONE autograd node -- what the
    training loop's ``sum(total_variation_loss(embeddings[i], ...) for i in range(n_levels))`` (run_nerf.py:628-635)
    adds up, without the 16 select / add nodes each way (CUDA-graph steps: every node costs a launch)."""
    key = (embed_fn.base_resolution, embed_fn.finest_resolution, embed_fn.log2_hashmap_size, embed_fn.n_levels)
    return _sweep_terms(embed_fn, key, embed_fn.base_resolution, embed_fn.finest_resolution,
                        embed_fn.log2_hashmap_size, embed_fn.n_levels)


def total_variation_loss(embeddings, min_resolution, max_resolution, level, log2_hashmap_size, n_levels=16):
    """Squared finite differences of one level's features over a random cube of grid vertices."""
    weight = embeddings._parameters['weight'] if 'weight' in embeddings._parameters else embeddings.weight
    owner, own_level = level_owner(embeddings)

    if (TV_SWEEP and owner is not None and weight.is_cuda and own_level == level
            and owner.n_levels == n_levels and owner.log2_hashmap_size == log2_hashmap_size):
        # a sweep is only reused for the same arguments and grad mode, and a level's term only while that level's
        # table is unchanged (storage, torch version counter, and the update counter of this package's RAdam, whose
        # kernels write through raw pointers)
        args = (min_resolution, max_resolution, log2_hashmap_size, n_levels, torch.is_grad_enabled())
        sw = _SWEEPS.get(owner)
        if (sw is None or sw.key != args or level <= sw.last or sw.epoch != ops.param_epoch[0]
                or sw.versions[level] != weight._version or sw.ptrs[level] != weight.data_ptr()):
            sw = _Sweep()
            sw.key = args
            sw.epoch = ops.param_epoch[0]
            sw.vec = _sweep_terms(owner, args, min_resolution, max_resolution, log2_hashmap_size, n_levels, parts=True)
            ws = owner._level_weights()
            sw.versions = [w._version for w in ws]
            sw.ptrs = [w.data_ptr() for w in ws]
            _SWEEPS[owner] = sw
        sw.last = level
        return sw.vec[level]

    resolution, cube = _level_cube(min_resolution, max_resolution, level, n_levels)
    # loss.py:25 draws on the default device; the reference only works when that is the tables' device, so draw
    # there directly (no host round trip)
    origin = torch.randint(0, resolution - cube, (3,), device=weight.device)
    if origin.device != weight.device:
        origin = origin.to(weight.device)
    sink_info = None
    if owner is not None and torch.is_grad_enabled() and weight.requires_grad:
        sink = owner.grad_sink()
        if sink is not None:
            sink_info = (sink, own_level * weight.numel())
    return ops.TVLossFn.apply(weight, origin, cube, log2_hashmap_size, sink_info)


def sigma_sparsity_loss(sigmas):
    """Cauchy sparsity prior on densities (unused by the training loop, kept for the import surface)."""
    return torch.log(1.0 + 2 * sigmas ** 2).sum(dim=-1)
