"""Experiment: how much do spatially coherent (sorted) points change the hash-encode kernels?"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
from hn_b200 import _lib, ops
from sweep_hash import timeit

def morton3(ix, iy, iz, bits):
    def spread(v):
        out = torch.zeros_like(v)
        for b in range(bits):
            out |= ((v >> b) & 1) << (3 * b)
        return out
    return spread(ix) | (spread(iy) << 1) | (spread(iz) << 2)

def main():
    n = int(os.environ.get("N", 1 << 24))
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
    dy = torch.randn(n, 32, device=dev, generator=gen)
    box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
    res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
    log2T = 19
    tables = (torch.rand(16 << log2T, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
    dt = torch.zeros_like(tables)
    def cell(r):
        return ((x + 1.5) / 3.0 * r).floor().clamp(0, r - 1).long()
    orders = {"random": None}
    c = cell(128); orders["linear128"] = torch.argsort(c[:, 0] + 128 * (c[:, 1] + 128 * c[:, 2]))
    c = cell(512); orders["linear512"] = torch.argsort(c[:, 0] + 512 * (c[:, 1] + 512 * c[:, 2]))
    orders["morton512"] = torch.argsort(morton3(c[:, 0], c[:, 1], c[:, 2], 9))
    c = cell(64); orders["linear64"] = torch.argsort(c[:, 0] + 64 * (c[:, 1] + 64 * c[:, 2]))
    t0 = timeit(lambda: torch.argsort(c[:, 0] + 64 * (c[:, 1] + 64 * c[:, 2])), 3)
    print(json.dumps({"torch_argsort_ms": round(t0, 3)}), flush=True)
    for name, perm in orders.items():
        xs = x if perm is None else x[perm].contiguous()
        for lpg in (2, 4, 16):
            _lib.set_tuning("hash_fwd_lpg", lpg); _lib.set_tuning("hash_bwd_lpg", lpg)
            tf = timeit(lambda: ops.hash_encode_forward(xs, tables, box, res, 16, 2, log2T, want_keep=False), 5)
            tb = timeit(lambda: ops.hash_encode_backward(xs, dy, box, res, 16, 2, log2T, dt), 5)
            print(json.dumps(dict(order=name, lpg=lpg, fwd_ms=round(tf, 3), bwd_ms=round(tb, 3))), flush=True)

if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
