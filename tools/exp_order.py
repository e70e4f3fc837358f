"""Which processing order suits the gather/scatter kernels?  Points are ordered in torch by different cell keys on a
256^3 grid (x-fastest rows = what hn_hash_sort_points produces; rows of b x b x b blocks; Morton) and the sorted
kernels are timed on the result."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import ops
from sweep_hash import timeit
n = 1 << 24; G = 256; dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
dy = torch.randn(n, 32, device=dev, generator=gen)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
tables = (torch.rand(16 << 19, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
dt = torch.zeros_like(tables)
c = ((x + 1.5) / 3.0 * G).floor().clamp(0, G - 1).long()
cx, cy, cz = c[:, 0], c[:, 1], c[:, 2]
def spread(v):  # Morton bit spread of 8-bit values
    v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
    return v
keys = {"rows_x_fastest": cx + G * (cy + G * cz)}
for b in ():
    nb = G // b
    keys[f"blocks_{b}"] = ((cx % b) + b * ((cy % b) + b * (cz % b))) + (b ** 3) * ((cx // b) + nb * ((cy // b) + nb * (cz // b)))
row = cy + G * cz
def bitrev16(v):
    r = torch.zeros_like(v)
    for i in range(16):
        r |= ((v >> i) & 1) << (15 - i)
    return r
keys["rows_scrambled"] = cx + G * ((row * 40503) % 65536)
f = ((x + 1.5) / 3.0 * (2 * G)).floor().clamp(0, 2 * G - 1).long()
octant = (f[:, 0] & 1) + 2 * (f[:, 1] & 1) + 4 * (f[:, 2] & 1)
base = cx + G * ((row * 40503) % 65536)
keys["rows_scrambled_then_octant"] = base * 8 + octant
keys["rows_scrambled_then_x_half"] = base * 2 + (f[:, 0] & 1)
for name, key in keys.items():
    order = torch.argsort(key, stable=True)
    xs4 = torch.cat([x[order], order.to(torch.int32).view(torch.float32)[:, None]], 1).contiguous()
    tf = timeit(lambda: ops.hash_encode_forward_sorted(xs4, tables, box, res, 16, 2, 19, want_keep=False), 10)
    tb = timeit(lambda: ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, 19, dt), 10)
    print(json.dumps(dict(order=name, fwd_ms=round(tf, 3), bwd_ms=round(tb, 3), sum=round(tf + tb, 3))), flush=True)
