"""Cost of splitting the sorted scatter into level buckets (no collective): 1 launch vs 2 / 4 launches."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import ops
from sweep_hash import timeit
n = 1 << 24; dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
dy = torch.randn(n, 32, device=dev, generator=gen)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
dt = torch.zeros(16 << 19, 2, device=dev)
xs4 = ops.hash_sort_points(x, box, 256)
for parts in ([(0, 16)], [(0, 8), (8, 16)], [(0, 4), (4, 8), (8, 12), (12, 16)]):
    def f():
        for b, e in parts:
            ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, 19, dt, levels=(b, e))
    print(json.dumps({"parts": len(parts), "bwd_ms": round(timeit(f, 10), 3)}), flush=True)
for b in range(0, 16, 2):
    f = lambda: ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, 19, dt, levels=(b, b + 2))
    print(json.dumps({"levels": [b, b + 2], "bwd_ms": round(timeit(f, 10), 3)}), flush=True)
