"""Regularisers -- drop-in for the reference's ``loss.py`` (total_variation_loss :11-43,
sigma_sparsity_loss :45-47).  Adjacent to the hot path ("next" row 2): the hashed
gather uses the CUDA spatial hash; the finite differences stay in PyTorch for now.
"""
from __future__ import annotations

from math import exp, floor, log

import torch

from embedding.hash_encoding import hash


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

    dev = embeddings.weight.device
    origin = torch.randint(0, resolution - cube, (3,)).to(dev)  # drawn like loss.py:25, then moved
    ax = torch.arange(cube + 1, device=dev)
    gx, gy, gz = torch.meshgrid(origin[0] + ax, origin[1] + ax, origin[2] + ax, indexing="ij")
    feats = embeddings(hash(torch.stack([gx, gy, gz], dim=-1), log2_hashmap_size))
    tv = torch.pow(feats[1:] - feats[:-1], 2).sum() \
        + torch.pow(feats[:, 1:] - feats[:, :-1], 2).sum() \
        + torch.pow(feats[:, :, 1:] - feats[:, :, :-1], 2).sum()
    return tv / cube


def sigma_sparsity_loss(sigmas):
    """Cauchy sparsity prior on densities (unused by the training loop, kept for the import surface)."""
    return torch.log(1.0 + 2 * sigmas ** 2).sum(dim=-1)
