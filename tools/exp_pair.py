"""Face exchange between x-adjacent runs in the aggregated scatter (hash_pair_min_heads / hash_pair_direct):
time and agreement with the plain aggregated scatter, on sorted uniform points and on ray-ordered samples.

The exchange was slower everywhere and its kernel code was not kept: this script needs the two tuning
keys of that experiment and is here as the record of what was measured, not as a runnable tool."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import ops, _lib
from sweep_hash import timeit
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
log2T = 19
n = 1 << 24
x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
dy = torch.randn(n, 32, device=dev, generator=gen)
xs4 = ops.hash_sort_points(x, box, 256)
# ray-ordered samples: 8192 rays x 256 samples through the box
R, S = 8192, 256
o = torch.tensor([0., 0., 4.], device=dev) + 0.1 * torch.randn(R, 3, device=dev, generator=gen)
d = -o / o.norm(dim=-1, keepdim=True) + 0.2 * torch.randn(R, 3, device=dev, generator=gen)
t = torch.sort(2. + 4. * torch.rand(R, S, device=dev, generator=gen), dim=-1).values
xr = (o[:, None, :] + d[:, None, :] * t[..., None]).reshape(-1, 3).contiguous()
dyr = torch.randn(R * S, 32, device=dev, generator=gen)
flush = torch.empty(48 << 20, dtype=torch.float32, device=dev)

def run_sorted(dt):
    ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, dt)

def run_ordered(dt):
    ops.hash_encode_backward(xr, dyr, box, res, 16, 2, log2T, dt, ordered=True)

def time_ordered(dt):
    # L2 flushed between launches (192 MiB written), its time subtracted
    tf = timeit(lambda: flush.fill_(1.0), 10)
    def both():
        flush.fill_(1.0); run_ordered(dt)
    return timeit(both, 10) - tf

ref = {}
for pmh, pdir in ((33, 0), (2, 0), (4, 0), (6, 0), (8, 0), (12, 0), (6, 1), (2, 1), (33, 1)):
    _lib.set_tuning("hash_pair_min_heads", pmh); _lib.set_tuning("hash_pair_direct", pdir)
    out = {"pair_min_heads": pmh, "pair_direct": pdir}
    for name, fn, tm in (("sorted", run_sorted, None), ("ordered", run_ordered, time_ordered)):
        dt = torch.zeros(16 << log2T, 2, device=dev)
        fn(dt)
        if name not in ref:
            ref[name] = dt.clone()
        err = (dt - ref[name]).abs().max().item() / ref[name].abs().max().item()
        ms = timeit(lambda: fn(dt), 5) if tm is None else tm(dt)
        out[name + "_ms"] = round(ms, 4); out[name + "_relerr"] = float("%.2e" % err)
    print(json.dumps(out), flush=True)
