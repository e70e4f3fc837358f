"""Time of the counting sort alone (and a permutation check), per sort variant (hash_sort_two_level)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import ops, _lib
from sweep_hash import timeit
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
variants = [int(v) for v in os.environ.get("VARIANTS", "1,0").split(",")]
for logn in (24, 22, 20):
    n = 1 << logn
    x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
    g = ops.sort_grid_res(n)
    for var in variants:
        _lib.set_tuning("hash_sort_two_level", var)
        t = timeit(lambda: ops.hash_sort_points(x, box, g), 10)
        xs4 = ops.hash_sort_points(x, box, g)
        rows = xs4[:, 3].contiguous().view(torch.int32).long()
        ok = torch.equal(torch.sort(rows).values, torch.arange(n, device=dev)) and torch.equal(xs4[:, :3], x[rows])
        print(json.dumps({"n": n, "grid": g, "variant": var, "sort_ms": round(t, 4), "permutation_ok": bool(ok)}), flush=True)
# clustered input: every point in 1/64 of the box (bins overflow the shared-memory capacity and are ordered chunk-wise)
n = 1 << 22
x = torch.rand(n, 3, device=dev, generator=gen) * 0.75 - 1.5
for var in variants:
    _lib.set_tuning("hash_sort_two_level", var)
    t = timeit(lambda: ops.hash_sort_points(x, box, 128), 5)
    xs4 = ops.hash_sort_points(x, box, 128)
    rows = xs4[:, 3].contiguous().view(torch.int32).long()
    ok = torch.equal(torch.sort(rows).values, torch.arange(n, device=dev)) and torch.equal(xs4[:, :3], x[rows])
    print(json.dumps({"n": n, "clustered": True, "variant": var, "sort_ms": round(t, 4), "permutation_ok": bool(ok)}), flush=True)
