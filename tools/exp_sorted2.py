"""Experiment: in-library counting sort + sorted forward + warp-aggregated backward."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib, ops
from sweep_hash import timeit

def main():
    n = int(os.environ.get("N", 1 << 24))
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
    dy = torch.randn(n, 32, device=dev, generator=gen)
    box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
    res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
    for log2T in (19,):
        tables = (torch.rand(16 << log2T, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
        dt = torch.zeros_like(tables)
        # reference results from the plain path
        for lpg in (4, 8):
            _lib.set_tuning("hash_fwd_lpg", lpg); _lib.set_tuning("hash_bwd_lpg", lpg)
            tf = timeit(lambda: ops.hash_encode_forward(x, tables, box, res, 16, 2, log2T, want_keep=False), 5)
            tb = timeit(lambda: ops.hash_encode_backward(x, dy, box, res, 16, 2, log2T, dt), 5)
            print(json.dumps(dict(mode="plain", lpg=lpg, fwd_ms=round(tf, 3), bwd_ms=round(tb, 3))), flush=True)
        out_ref, _ = ops.hash_encode_forward(x, tables, box, res, 16, 2, log2T)
        dt.zero_(); ops.hash_encode_backward(x, dy, box, res, 16, 2, log2T, dt); dt_ref = dt.clone()
        for grid in (128, 256):
            ts = timeit(lambda: ops.hash_sort_points(x, box, grid), 5)
            xs4 = ops.hash_sort_points(x, box, grid)
            rows = xs4[:, 3].contiguous().view(torch.int32).long()
            assert torch.equal(torch.sort(rows).values, torch.arange(n, device=dev)), "not a permutation"
            for lpg, lm in ((4, 0), (4, 1), (8, 0), (8, 1), (16, 0)):
                _lib.set_tuning("hash_fwd_lpg", lpg); _lib.set_tuning("hash_bwd_lpg", lpg)
                _lib.set_tuning("hash_level_major", lm)
                tf = timeit(lambda: ops.hash_encode_forward_sorted(xs4, tables, box, res, 16, 2, log2T, want_keep=False), 5)
                tb = timeit(lambda: ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, dt), 5)
                out, _ = ops.hash_encode_forward_sorted(xs4, tables, box, res, 16, 2, log2T)
                same = torch.equal(out, out_ref)
                dt.zero_(); ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, dt)
                err = ((dt - dt_ref).abs().max() / dt_ref.abs().max()).item()
                print(json.dumps(dict(log2T=log2T, grid=grid, lpg=lpg, lm=lm, sort_ms=round(ts, 3), fwd_ms=round(tf, 3),
                                      bwd_ms=round(tb, 3), total_ms=round(ts + tf + tb, 3), fwd_bit_equal=same,
                                      grad_rel_err=err)), flush=True)
        del tables, dt, dt_ref, out_ref

if __name__ == "__main__":
    main()
