"""Does a finer sort grid pay?  fwd/bwd time on points sorted with grid 256 vs 384 vs 512 (single-pass sort used for
the grids the two-level sort does not cover; only the encode kernels are timed)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib, ops
from sweep_hash import timeit
n = 1 << 24; dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
dy = torch.randn(n, 32, device=dev, generator=gen)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
for log2T in (19,):
    tables = (torch.rand(16 << log2T, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
    dt = torch.zeros_like(tables)
    for grid in (128, 160, 192, 224, 256):
        _lib.set_tuning("hash_sort_two_level", 1 if grid <= 256 else 0)
        ts = timeit(lambda: ops.hash_sort_points(x, box, grid), 3)
        xs4 = ops.hash_sort_points(x, box, grid)
        tf = timeit(lambda: ops.hash_encode_forward_sorted(xs4, tables, box, res, 16, 2, log2T, want_keep=False), 10)
        tb = timeit(lambda: ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, dt), 10)
        print(json.dumps(dict(log2T=log2T, grid=grid, sort_ms=round(ts, 3), fwd_ms=round(tf, 3), bwd_ms=round(tb, 3))), flush=True)
    del tables, dt
_lib.set_tuning("hash_sort_two_level", 1)
