"""Deterministic synthetic inputs shared by the golden generator, the tests and the bench.

TEST INFRASTRUCTURE (see oracle/oracle.py header).  Everything is derived from integer formulas
or ``numpy.random.RandomState`` (whose stream is frozen by numpy's compatibility policy), so the
inputs can be regenerated bit-identically on any box without storing them.
"""
from __future__ import annotations

import numpy as np

# the power-of-two-friendly micro-benchmark box and the rounding-hostile parity box.
BBOX_UNIT = ((-1.5, -1.5, -1.5), (1.5, 1.5, 1.5))
BBOX_ODD = ((-4.6371, -4.4123, -3.0517), (4.5519, 4.7311, 5.0313))


def synth_tables(n_levels: int, log2_hashmap_size: int, n_features: int) -> np.ndarray:
    """[L, 2^T, F] fp32 in [-1e-4, 1e-4) (the reference's init range, hash_encoding.py:56),
    from a multiplicative integer hash so no RNG implementation is involved."""
    T = 1 << log2_hashmap_size
    lvl = np.arange(n_levels, dtype=np.uint64)[:, None, None]
    row = np.arange(T, dtype=np.uint64)[None, :, None]
    col = np.arange(n_features, dtype=np.uint64)[None, None, :]
    h = (row * np.uint64(2654435761) + lvl * np.uint64(40503) + col * np.uint64(9973) + np.uint64(12345))
    h = (h ^ (h >> np.uint64(13))) * np.uint64(1274126177)
    h = (h ^ (h >> np.uint64(16))) & np.uint64(0xFFFF)
    return ((h.astype(np.float64) / 65536.0) * 2e-4 - 1e-4).astype(np.float32)


def points_in_box(n: int, bbox, seed: int, adversarial: bool = True) -> np.ndarray:
    """[n,3] fp32 uniform in the box; with ``adversarial`` the first rows are replaced by points
    exactly on box_min / box_max, on cell boundaries, and outside the box."""
    lo = np.asarray(bbox[0], dtype=np.float32)
    hi = np.asarray(bbox[1], dtype=np.float32)
    rs = np.random.RandomState(seed)
    x = (lo + (hi - lo) * rs.rand(n, 3).astype(np.float32)).astype(np.float32)
    x = np.minimum(np.maximum(x, lo), hi)
    if adversarial and n >= 16:
        x[0] = lo
        x[1] = hi
        x[2] = (lo[0], hi[1], lo[2])
        x[3] = (hi[0], lo[1], hi[2])
        x[4] = lo - np.float32(0.75)            # outside: clamped for indices, extrapolated weights
        x[5] = hi + np.float32(1.25)
        x[6] = (lo[0] - 2.0, 0.5 * (lo[1] + hi[1]), hi[2] + 0.5)
        x[7] = 0.5 * (lo + hi)                  # centre
        span = hi - lo
        x[8] = lo + span * np.float32(0.25)     # cell boundaries at power-of-two resolutions
        x[9] = lo + span * np.float32(0.5)
        x[10] = lo + span * np.float32(1.0 / 16.0)
        x[11] = np.nextafter(hi, lo).astype(np.float32)
        x[12] = np.nextafter(lo, hi).astype(np.float32)
        x[13] = lo + span * np.float32(31.0 / 32.0)
        x[14] = (hi[0], hi[1], lo[2])
        x[15] = lo + span * np.float32(1.0 / 512.0)
    return x.astype(np.float32)


def mlp_weights(seed: int, input_ch: int = 32, input_ch_views: int = 16, hidden: int = 64,
                geo: int = 15, scale: float = 1.0):
    """NeRFSmall weights as instantiated at run_nerf_helpers.py:79-84, drawn U(-1/sqrt(in), 1/sqrt(in))
    (nn.Linear's default bound) from RandomState so they do not depend on torch's RNG."""
    rs = np.random.RandomState(seed)

    def lin(o, i):
        b = scale / np.sqrt(i)
        return ((rs.rand(o, i) * 2 - 1) * b).astype(np.float32)

    sigma = [lin(hidden, input_ch), lin(1 + geo, hidden)]
    color = [lin(hidden, input_ch_views + geo), lin(hidden, hidden), lin(3, hidden)]
    return sigma, color


def rays(n_rays: int, seed: int, near: float = 2.0, far: float = 6.0) -> np.ndarray:
    """[R, 11] ray batch (o, d, near, far, viewdir) per SURVEY 8d cfg1: origins ~N((0,0,4), 0.1^2),
    directions toward the scene origin + 0.2 N(0,1), unnormalised; viewdir = d/|d|."""
    rs = np.random.RandomState(seed)
    o = (np.array([0.0, 0.0, 4.0]) + 0.1 * rs.randn(n_rays, 3)).astype(np.float32)
    d = (-o / np.linalg.norm(o, axis=-1, keepdims=True) + 0.2 * rs.randn(n_rays, 3)).astype(np.float32)
    v = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    nf = np.tile(np.array([[near, far]], dtype=np.float32), (n_rays, 1))
    return np.concatenate([o, d, nf, v], axis=-1).astype(np.float32)
