"""Import-surface shim for the reference's ``embedding/embedder.py``.

The frequency (sin/cos) encoder belongs to the ``i_embed=0`` branch, which the reference's own
``run_network`` cannot execute (it unpacks a tuple the frequency encoder does not return --
run_nerf_helpers.py:216) and which is outside the hash-encoding hot path.  The class
is kept, in plain torch, only so that ``from embedding.embedder import get_embedder, Embedder``
(run_nerf_helpers.py:19) resolves; ``get_embedder`` lives in run_nerf_helpers like in the reference and is
re-exported lazily to avoid a circular import.
"""
from __future__ import annotations

import torch


class Embedder:
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs["input_dims"]
        n = kwargs["num_freqs"]
        top = kwargs["max_freq_log2"]
        if kwargs.get("log_sampling", True):
            self.freq_bands = 2.0 ** torch.linspace(0.0, top, steps=n)
        else:
            self.freq_bands = torch.linspace(1.0, 2.0 ** top, steps=n)
        self.periodic_fns = list(kwargs.get("periodic_fns", (torch.sin, torch.cos)))
        self.include_input = bool(kwargs.get("include_input", True))
        self.out_dim = d * (int(self.include_input) + n * len(self.periodic_fns))

    def embed(self, inputs):
        parts = [inputs] if self.include_input else []
        for f in self.freq_bands:
            parts.extend(fn(inputs * f) for fn in self.periodic_fns)
        return torch.cat(parts, dim=-1)


def get_embedder(multires, args, i=0):
    from run_nerf_helpers import get_embedder as _impl
    return _impl(multires, args, i)
