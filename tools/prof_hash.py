"""One pass of every hash-encode kernel at bench size, for ncu captures (no timing here)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
from hn_b200 import _lib, ops

n = int(os.environ.get("N", 1 << 24))
log2T = int(os.environ.get("LOG2T", 19))
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
dy = torch.randn(n, 32, device=dev, generator=gen)
box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
tables = (torch.rand(16 << log2T, 2, device=dev, generator=gen) * 2e-4 - 1e-4)
dt = torch.zeros_like(tables)
torch.cuda.synchronize()
xs4 = ops.hash_sort_points(x, box, int(os.environ.get("GRID", 256)))
ops.hash_encode_forward_sorted(xs4, tables, box, res, 16, 2, log2T)
ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, dt)
ops.hash_encode_forward(x, tables, box, res, 16, 2, log2T)
ops.hash_encode_backward(x, dy, box, res, 16, 2, log2T, dt)
torch.cuda.synchronize()
print("ok")
