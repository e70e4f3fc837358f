"""CUDA-event timing of the MLP forward and backward (both backward implementations) at training size."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
from models import NeRFSmall
from hn_b200 import _lib
dev = torch.device("cuda:0")
net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16).to(dev)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = R * 192
e = (torch.randn(N, 32, device=dev) * 0.3).requires_grad_(True)
v = torch.randn(R, 16, device=dev)
dO = torch.randn(N, 4, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for impl in (1, 0):
    _lib.set_tuning("mlp_bwd_impl", impl)
    for _ in range(3):
        out = net.forward_fused(e, v, 192, None); out.backward(dO)
    torch.cuda.synchronize()
    with torch.no_grad():
        t_f = timed(lambda: net.forward_fused(e, v, 192, None))
    def fb():
        net.forward_fused(e, v, 192, None).backward(dO)
    t_fb = timed(fb)
    print(f"mlp_bwd_impl={impl} N={N}: forward {t_f:.3f} ms, forward+backward {t_fb:.3f} ms, backward {t_fb - t_f:.3f} ms", flush=True)
_lib.set_tuning("mlp_bwd_impl", 1)
