import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib, ops
from models import NeRFSmall
from sweep_hash import timeit
dev = torch.device("cuda:0")
net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16).to(dev)
N = 8192 * 192
e = (torch.randn(N, 32, device=dev) * 0.3).requires_grad_(True); v = torch.randn(8192, 16, device=dev); dO = torch.randn(N, 4, device=dev)
def fb():
    out = net.forward_fused(e, v, 192, None); out.backward(dO)
for ab in (0, 1, 2, 4, 3, 5, 6, 7):
    _lib.set_tuning("mlp_dw_ablate", ab)
    print(json.dumps({"ablate": ab, "fwd_bwd_ms": round(timeit(fb, 5), 3)}), flush=True)
