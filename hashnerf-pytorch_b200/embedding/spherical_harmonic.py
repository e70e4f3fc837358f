"""Spherical-harmonics direction encoding -- drop-in for the reference's ``embedding/spherical_harmonic.py``."""
from __future__ import annotations

import torch.nn as nn

from hn_b200 import ops


class SHEncoder(nn.Module):
    """Real SH basis up to degree 5 (reference spherical_harmonic.py:43-103); one fused kernel instead of
    ~60 strided elementwise launches.  No trainable state."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        assert self.input_dim == 3
        assert 1 <= self.degree <= 5
        self.out_dim = degree ** 2

    def forward(self, input, **kwargs):
        return ops.sh_encode(input, self.degree)
