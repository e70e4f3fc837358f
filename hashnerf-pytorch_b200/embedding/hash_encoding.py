"""Multiresolution hash encoding -- drop-in for the reference's ``embedding/hash_encoding.py``.

Same public surface (``HashEmbedder``, ``hash``, ``HASH_PRIMES``, ``BOX_OFFSETS`` and, because the
reference's helper module looks for it here -- run_nerf_helpers.py:20, SURVEY Appendix B2 --
``SHEncoder``), but the per-level chain of ~70 ATen ops (hash_encoding.py:59-163) is ONE CUDA kernel
launch for all levels and its autograd is ONE scatter kernel, through libhashnerf_b200.so.

CUDA only; a CPU tensor raises RuntimeError.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from hn_b200 import ops
from .spherical_harmonic import SHEncoder  # noqa: F401  (re-export, see module docstring)

# hash_encoding.py:7
HASH_PRIMES = [1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737]


def __getattr__(name):
    # hash_encoding.py:10-11 allocates this on 'cuda' at import time (Appendix B4); build it on demand
    # on the default device instead so the module imports without a driver.
    if name == "BOX_OFFSETS":
        return torch.tensor([[[(c >> 2) & 1, (c >> 1) & 1, c & 1] for c in range(8)]])
    raise AttributeError(name)


def hash(coords, log2_hashmap_size):
    """Spatial hash of integer coordinates (reference hash_encoding.py:112-128).

    coords: integer tensor [..., dim], dim <= 7 -> int64 [...] in [0, 2**log2_hashmap_size)."""
    return ops.spatial_hash(coords, log2_hashmap_size)


_LEVEL_OWNER = weakref.WeakKeyDictionary()  # nn.Embedding of a level -> (weakref to its HashEmbedder, level index)


def level_owner(embedding_module):
    """(HashEmbedder, level) the module is level ``level`` of, or (None, None)."""
    entry = _LEVEL_OWNER.get(embedding_module)
    if entry is None:
        return None, None
    owner = entry[0]()
    # plain dict lookups (nn.ModuleList.__getitem__ is ~2 us, and the training loop asks 16 times per step)
    if owner is None or owner._modules['embeddings']._modules.get(str(entry[1])) is not embedding_module:
        return None, None
    return owner, entry[1]


class HashEmbedder(nn.Module):
    """Reference hash_encoding.py:13-110.

    Parameters keep the reference's names and shapes (``embeddings.{l}.weight`` of shape
    [2**log2_hashmap_size, n_features_per_level]) so checkpoints interchange, but they are views of one
    flat ``[L, 2^T, F]`` buffer: the kernels see a single table and the optimizer/all-reduce can treat the
    encoder as one tensor."""

    def __init__(self, bounding_box, n_levels=16, n_features_per_level=2,
                 log2_hashmap_size=19, base_resolution=16, finest_resolution=512):
        super().__init__()
        if n_features_per_level not in (1, 2, 4):
            raise ValueError("n_features_per_level must be 1, 2 or 4 (the reference uses 2)")
        if not 1 <= n_levels <= 32:
            raise ValueError("n_levels must be in [1, 32]")
        self.bounding_box = bounding_box
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = torch.tensor(base_resolution)
        self.finest_resolution = torch.tensor(finest_resolution)
        self.out_dim = n_levels * n_features_per_level
        # growth factor, same fp32 tensor ops as hash_encoding.py:50 (device = wherever tensors default to)
        self.b = torch.exp((torch.log(self.finest_resolution) - torch.log(self.base_resolution)) / (n_levels - 1)) \
            if n_levels > 1 else torch.tensor(float("nan"))

        # Same construction order as hash_encoding.py:52-56 so a seeded RNG yields the same tables.
        rows = 2 ** log2_hashmap_size
        self.embeddings = nn.ModuleList([nn.Embedding(rows, n_features_per_level) for _ in range(n_levels)])
        for emb in self.embeddings:
            nn.init.uniform_(emb.weight, a=-0.0001, b=0.0001)
        self._flatten_parameters()
        self._geom_cache = {}
        # let code that is handed a single level (loss.total_variation_loss gets embeddings[i]) find its way back
        # to the shared gradient buffer
        self._register_levels()
        # None: re-order large batches by grid cell before encoding (results unchanged, see ops.HashEncodeFn);
        # True / False force the choice.
        self.coherent = None
        # True: backward accumulates in place into one persistent flat gradient buffer and points
        # embeddings[l].weight.grad at its slices (see ops.GradSink) -- what loss.backward() + optimizer.step() see is
        # identical to autograd's, but the autograd node itself returns no parameter gradient, so
        # torch.autograd.grad(loss, params), parameter hooks and DDP need False: plain autograd gradients.
        self.fused_grad_accumulation = True
        self._sink = None

    # -- level modules -> encoder ------------------------------------------------------------------
    def _register_levels(self):
        """Let the per-level modules find their encoder (total_variation_loss is handed a bare
        ``embed_fn.embeddings[i]``).  Kept in a module-level weak registry, not on the modules: nothing extra is
        pickled or deep-copied, and a copied encoder registers its own children (see __setstate__)."""
        ref = weakref.ref(self)
        for l, emb in enumerate(self.embeddings):
            _LEVEL_OWNER[emb] = (ref, l)

    def __setstate__(self, state):
        super().__setstate__(state)
        self._sink = None          # gradient buffers belong to the object they were created for
        self._flat_ok = False      # copied level tables are separate tensors until the next flatten
        self._register_levels()

    # -- storage --------------------------------------------------------------------------------
    def _level_weights(self):
        # plain dict lookups: nn.Module.__getattr__ on 16 children per call is measurable in the eager step
        return [m._parameters['weight'] for m in self._modules['embeddings']._modules.values()]

    def _flatten_parameters(self):
        """Re-home the level tables into one contiguous buffer (keeps Parameter identity)."""
        ws = self._level_weights()
        # cheap steady-state check: the end points of the span are where a flat buffer puts them
        if getattr(self, "_flat_ok", False) and \
                ws[-1].data_ptr() - ws[0].data_ptr() == (len(ws) - 1) * ws[0].numel() * 4:
            return
        if ops._consecutive(ws):
            self._flat_ok = True
            return
        self._flat_ok = True
        flat = torch.empty(len(ws), *ws[0].shape, dtype=torch.float32, device=ws[0].device)
        with torch.no_grad():
            for l, w in enumerate(ws):
                flat[l].copy_(w)
                w.data = flat[l]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)  # .to()/.cuda() move each level separately
        self._flat_ok = False
        self._flatten_parameters()
        self._geom_cache = {}
        return out

    def flat_tables(self) -> torch.Tensor:
        """[L, 2^T, F] view over all level tables (no copy)."""
        self._flatten_parameters()
        w0 = self.embeddings[0].weight
        return torch.as_strided(w0.detach(), (self.n_levels,) + tuple(w0.shape),
                                (w0.numel(), w0.stride(0), w0.stride(1)))

    # -- geometry -------------------------------------------------------------------------------
    def level_resolutions(self) -> torch.Tensor:
        """floor(base * b**i) for every level with the reference's own tensor ops (hash_encoding.py:101);
        evaluated wherever ``b`` lives, so the integers match what the reference would compute there."""
        return torch.stack([torch.floor(self.base_resolution * self.b ** i) for i in range(self.n_levels)]) \
            .to(torch.float32)

    def _geometry(self, device):
        key = (device.type, device.index)
        g = self._geom_cache.get(key)
        if g is None:
            lo, hi = self.bounding_box
            box = torch.cat([torch.as_tensor(lo).reshape(3).float().to(device),
                             torch.as_tensor(hi).reshape(3).float().to(device)]).contiguous()
            g = (box, self.level_resolutions().to(device).contiguous())
            self._geom_cache[key] = g
        return g

    # -- forward --------------------------------------------------------------------------------
    def forward(self, x):
        """x: [N,3] points -> (features [N, n_levels*F], keep_mask [N] bool)  (hash_encoding.py:84-110)."""
        feats, keep = self.encode(x)
        return feats, keep.bool()

    def grad_sink(self):
        """The persistent flat gradient buffer manager of the level tables (None if disabled)."""
        if not self.fused_grad_accumulation:
            return None
        if self._sink is None or any(a is not b for a, b in zip(self._sink.params, self._level_weights())):
            self._sink = ops.GradSink(self._level_weights())
        return self._sink

    def encode(self, x, ordered=False):
        """Same as forward() but the mask stays the kernel's uint8 (what the fused MLP consumes).  ``ordered``:
        the caller vouches that consecutive points are spatial neighbours (samples along rays, as run_network
        produces them): sorting such points only costs (measured 800x800 frame: 159 ms sorted vs 69 ms as given)."""
        lead = x.shape[:-1]
        pts = x.reshape(-1, 3)
        box, res = self._geometry(pts.device)
        self._flatten_parameters()
        sink = self.grad_sink() if torch.is_grad_enabled() else None
        feats, keep = ops.HashEncodeFn.apply(pts, box, res, self.log2_hashmap_size, self.n_features_per_level,
                                             "ordered" if (ordered and self.coherent is None) else self.coherent, sink,
                                             *self._level_weights())
        if len(lead) != 1:
            feats, keep = feats.reshape(*lead, self.out_dim), keep.reshape(lead)
        return feats, keep

    def voxel_vertices(self, x):
        """Per-level (hashed [L,N,8] int64, vmin [L,N,3], vmax [L,N,3]) -- what get_voxel_vertices
        (hash_encoding.py:59-82) returns level by level; exposed for parity tests."""
        pts = x.reshape(-1, 3)
        box, res = self._geometry(pts.device)
        return ops.voxel_vertices(pts, box, res, self.log2_hashmap_size)
