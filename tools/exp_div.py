"""A/B of the hoisted-reciprocal index division (hn_set_tuning hash_div_hoist) on the sorted and plain paths."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib, ops
from sweep_hash import timeit

n = 1 << 24
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
dy = torch.randn(n, 32, device=dev, generator=gen)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
log2T = 19
tables = (torch.rand(16 << log2T, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
dt = torch.zeros_like(tables)
xs4 = ops.hash_sort_points(x, box, 256)
for rep in range(3):
    for hoist in (0, 1):
        _lib.set_tuning("hash_div_hoist", hoist)
        tf = timeit(lambda: ops.hash_encode_forward_sorted(xs4, tables, box, res, 16, 2, log2T, want_keep=False), 10)
        tb = timeit(lambda: ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, dt), 10)
        pf = timeit(lambda: ops.hash_encode_forward(x, tables, box, res, 16, 2, log2T, want_keep=False), 5)
        print(json.dumps(dict(hoist=hoist, fwd_sorted_ms=round(tf, 3), bwd_sorted_ms=round(tb, 3), fwd_plain_ms=round(pf, 3))), flush=True)
