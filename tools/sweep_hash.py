"""GPU sweep of the hash-encode launch shapes (levels per thread) across table sizes.  Writes one JSON
line per configuration; used to pick heuristics, not a bench."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
from hn_b200 import _lib, ops  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(os.environ.get("N", 1 << 24))
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
    dy = torch.randn(n, 32, device=dev, generator=gen)
    box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
    res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
    for log2T in (14, 19, 22):
        tables = (torch.rand(16 << log2T, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
        dt = torch.zeros_like(tables)
        for lpg in (1, 2, 4, 8, 16):
            _lib.set_tuning("hash_fwd_lpg", lpg)
            _lib.set_tuning("hash_bwd_lpg", lpg)
            tf = timeit(lambda: ops.hash_encode_forward(x, tables, box, res, 16, 2, log2T, want_keep=False), 5)
            tb = timeit(lambda: ops.hash_encode_backward(x, dy, box, res, 16, 2, log2T, dt), 5)
            print(json.dumps(dict(log2T=log2T, lpg=lpg, n=n, fwd_ms=round(tf, 3), bwd_ms=round(tb, 3),
                                  fwd_gbs=round(n * 1164 / tf / 1e6, 1), bwd_gbs=round(n * 1164 / tb / 1e6, 1),
                                  msamples_fwd_bwd=round(n / (tf + tb) / 1e3, 1))), flush=True)
        del tables, dt


if __name__ == "__main__":
    main()
