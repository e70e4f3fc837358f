"""Sort-grid choice at smaller batches: cube-root rule vs neighbouring powers of two (sort + fwd + bwd)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib, ops
from sweep_hash import timeit
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
tables = (torch.rand(16 << 19, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
dt = torch.zeros_like(tables)
for logn in (22, 20):
    n = 1 << logn
    x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
    dy = torch.randn(n, 32, device=dev, generator=gen)
    rule = ops.sort_grid_res(n)
    for grid in sorted({64, 128, 256, rule}):
        ts = timeit(lambda: ops.hash_sort_points(x, box, grid), 5)
        xs4 = ops.hash_sort_points(x, box, grid)
        tf = timeit(lambda: ops.hash_encode_forward_sorted(xs4, tables, box, res, 16, 2, 19, want_keep=False), 10)
        tb = timeit(lambda: ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, 19, dt), 10)
        print(json.dumps(dict(n=n, grid=grid, rule=(grid == rule), sort_ms=round(ts, 3), fwd_ms=round(tf, 3),
                              bwd_ms=round(tb, 3), total=round(ts + tf + tb, 3))), flush=True)
