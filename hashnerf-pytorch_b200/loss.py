"""Regularisers -- drop-in for the reference's ``loss.py`` (total_variation_loss :11-43,
sigma_sparsity_loss :45-47).  Adjacent to the hot path ("next" row 2): one CUDA launch for
the forward (hash, gather, squared differences, reduction) and one for the backward, instead of ~15 ATen
launches per level.
"""
from __future__ import annotations

from math import exp, floor, log

import torch

from embedding.hash_encoding import hash  # noqa: F401  (reference import line loss.py:8)
from hn_b200 import ops


def total_variation_loss(embeddings, min_resolution, max_resolution, level, log2_hashmap_size, n_levels=16):
    """Squared finite differences of one level's features over a random cube of grid vertices.

    NOTE the resolution here is fp64 ``math`` arithmetic (loss.py:13-14) while the encoder's is fp32 tensor
    arithmetic; they can disagree and both are kept as the reference has them."""
    growth = exp((log(max_resolution) - log(min_resolution)) / (n_levels - 1))
    resolution = int(floor(min_resolution * growth ** level))
    smallest = int(min_resolution) - 1
    largest = 50
    if smallest > largest:
        raise ValueError("ALERT! min cuboid size greater than max!")  # the reference drops into pdb here
    cube = int(floor(min(max(resolution / 10.0, smallest), largest)))

    weight = embeddings.weight
    # loss.py:25 draws on the default device; the reference only works when that is the tables' device, so draw
    # there directly (no host round trip)
    origin = torch.randint(0, resolution - cube, (3,), device=weight.device)
    if origin.device != weight.device:
        origin = origin.to(weight.device)
    sink_info = None
    owner = getattr(embeddings, "_hn_owner", None)
    owner = owner() if owner is not None else None
    if owner is not None and torch.is_grad_enabled() and weight.requires_grad:
        sink = owner.grad_sink()
        if sink is not None:
            sink_info = (sink, embeddings._hn_level * weight.numel())
    return ops.TVLossFn.apply(weight, origin, cube, log2_hashmap_size, sink_info)


def sigma_sparsity_loss(sigmas):
    """Cauchy sparsity prior on densities (unused by the training loop, kept for the import surface)."""
    return torch.log(1.0 + 2 * sigmas ** 2).sum(dim=-1)
