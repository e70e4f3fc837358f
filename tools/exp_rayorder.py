"""Inference-order experiment: hash forward on ray-major [R,S,3] points vs the same points sample-major [S,R,3]
(lanes = adjacent pixels' rays at the same sample index)."""
import json, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import ops
from ray_util import get_rays
from sweep_hash import timeit
dev = torch.device("cuda:0")
H = W = 800
focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
c2w = torch.tensor([[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 4.0]], device=dev)
o, d = get_rays(H, W, K, c2w)
o, d = o.reshape(-1, 3)[:32768 * 4], d.reshape(-1, 3)[:32768 * 4]      # 4 chunks of image rows
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
tables = (torch.rand(16 << 19, 2, device=dev) * 2e-4 - 1e-4)
for S in (64, 192):
    z = torch.linspace(2., 6., S, device=dev)[None, :] + 0.0 * o[:, :1]
    z = z + (torch.rand_like(z) - 0.5) * (4.0 / S)                      # jittered depths, sorted per ray
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).contiguous()    # [R,S,3]
    ray_major = pts.reshape(-1, 3)
    sample_major = pts.transpose(0, 1).contiguous().reshape(-1, 3)      # [S,R,3]
    # tiles of 32 rays x 8 samples, as a remapped kernel would walk them
    R = pts.shape[0]
    tiled = pts.reshape(R // 32, 32, S // 8, 8, 3).permute(0, 2, 3, 1, 4).contiguous().reshape(-1, 3)
    out = {}
    for name, p in (("ray_major", ray_major), ("sample_major", sample_major), ("tile_32rays_x_8samples", tiled)):
        for hint in (False,):
            t = timeit(lambda: ops.hash_encode_forward(p, tables, box, res, 16, 2, 19, want_keep=False), 5)
            out[name] = round(t, 3)
    print(json.dumps({"rays": R, "S": S, "points": R * S, **out}), flush=True)
