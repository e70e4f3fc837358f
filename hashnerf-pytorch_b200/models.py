"""Network modules -- drop-in for the part of the reference's ``models.py`` that is on the hash-NeRF path.

``NeRFSmall`` (reference models.py:96-174) keeps its constructor, parameter names
(``sigma_net.{i}.weight``, ``color_net.{i}.weight``) and ``forward(x[N, input_ch + input_ch_views])``
contract, but all five bias-free linears, the ReLUs and the slice/cat glue run as one fused CUDA kernel
(forward) and two (backward) from libhashnerf_b200.so.  ``NeRF`` / ``NeRFGradient`` (models.py:11-93,
177-212) belong to the ``i_embed=0`` branch, which is outside the hot path and cannot run in the
reference either; the names exist so the reference's import line resolves.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from hn_b200 import ops


class NeRFSmall(nn.Module):
    def __init__(self, num_layers=3, hidden_dim=64, geo_feat_dim=15, num_layers_color=4,
                 hidden_dim_color=64, input_ch=3, input_ch_views=3):
        super().__init__()
        self.input_ch = input_ch
        self.input_ch_views = input_ch_views
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.geo_feat_dim = geo_feat_dim
        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color

        # layer shapes exactly as models.py:116-147 (creation order matters for a seeded RNG)
        dims = [input_ch] + [hidden_dim] * (num_layers - 1) + [1 + geo_feat_dim]
        self.sigma_net = nn.ModuleList(nn.Linear(dims[i], dims[i + 1], bias=False) for i in range(num_layers))
        # NOTE models.py:139 uses hidden_dim (not hidden_dim_color) for the inner colour widths
        cdims = [input_ch_views + geo_feat_dim] + [hidden_dim] * (num_layers_color - 1) + [3]
        self.color_net = nn.ModuleList(nn.Linear(cdims[i], cdims[i + 1], bias=False) for i in range(num_layers_color))

        # The fused tcgen05 kernels implement the network create_nerf instantiates (run_nerf_helpers.py:79-84:
        # num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32, input_ch_views=16).  Any other
        # geometry -- including the constructor's own defaults, which nothing on the training path uses -- keeps the
        # signature's behaviour through the layer-by-layer form of models.py:151-174 on the same CUDA tensors (bias-free
        # F.linear = library SGEMM): correct, not a hot path, and still CUDA-only.
        got = tuple(tuple(l.weight.shape) for l in list(self.sigma_net) + list(self.color_net))
        self.fused = (got == ops.MLP_SHAPES)
        if self.fused:
            self._flatten_parameters()
        self.fused_grad_accumulation = True  # see ops.GradSink
        self._sink = None

    def _weights(self):
        return [l.weight for l in self.sigma_net] + [l.weight for l in self.color_net]

    def _flatten_parameters(self):
        ws = self._weights()
        if ops._consecutive(ws):
            return
        flat = torch.empty(ops.MLP_PARAMS, dtype=torch.float32, device=ws[0].device)
        off = 0
        with torch.no_grad():
            for w in ws:
                n = w.numel()
                flat[off:off + n].view_as(w).copy_(w)
                w.data = flat[off:off + n].view_as(w)
                off += n

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if self.fused:
            self._flatten_parameters()
        return out

    def _forward_layers(self, x):
        """models.py:151-174 for geometries the fused kernels do not cover."""
        if not x.is_cuda:
            raise RuntimeError("hashnerf_b200 runs on CUDA (sm_100a) only: there is no CPU fallback")
        inp, views = torch.split(x, [self.input_ch, self.input_ch_views], dim=-1)
        h = inp
        for l, lin in enumerate(self.sigma_net):
            h = torch.nn.functional.linear(h, lin.weight)
            if l != self.num_layers - 1:
                h = torch.relu(h)
        sigma, geo = h[..., 0], h[..., 1:]
        h = torch.cat([views, geo], dim=-1)
        for l, lin in enumerate(self.color_net):
            h = torch.nn.functional.linear(h, lin.weight)
            if l != self.num_layers_color - 1:
                h = torch.relu(h)
        return torch.cat([h, sigma.unsqueeze(dim=-1)], -1)

    def flat_weights(self) -> torch.Tensor:
        self._flatten_parameters()
        return ops.pack(self._weights())

    def forward(self, x):
        """x: [N, 48] = [hash features (32) | view features (16)] -> [N, 4] = [rgb_raw (3) | sigma (1)]."""
        if not self.fused:
            return self._forward_layers(x)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        out = self.forward_fused(x2[:, :self.input_ch], x2[:, self.input_ch:], 1, None)
        return out.reshape(*lead, 4)

    def forward_fused(self, enc, views, pts_per_view=1, keep=None):
        """enc [N,32]; views [ceil(N/pts_per_view),16] (one row per ray); keep [N] bool or None (sigma is
        zeroed where False, run_nerf_helpers.py:225)."""
        if not self.fused:
            raise NotImplementedError("forward_fused needs the geometry create_nerf instantiates (see __init__)")
        self._flatten_parameters()
        sink = None
        if self.fused_grad_accumulation and torch.is_grad_enabled():
            if self._sink is None or any(a is not b for a, b in zip(self._sink.params, self._weights())):
                self._sink = ops.GradSink(self._weights())
            sink = self._sink
        return ops.MLPFn.apply(enc, views, pts_per_view, keep, sink, *self._weights())


class _OutOfScope(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError(
            f"{type(self).__name__} is the positional-encoding (i_embed=0) network; this package implements "
            "the hash-encoding path only (i_embed=1 -> NeRFSmall)")


class NeRF(_OutOfScope):
    pass


class NeRFGradient(_OutOfScope):
    pass
