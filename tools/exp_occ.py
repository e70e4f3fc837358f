"""MLP forward: one vs two CTAs (tile contexts) per SM -- how much does a second context hide of the layer chain?"""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib
from models import NeRFSmall
from sweep_hash import timeit
dev = torch.device("cuda:0")
net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16).to(dev)
N = 8192 * 192
e = torch.randn(N, 32, device=dev) * 0.3; v = torch.randn(8192, 16, device=dev)
def f():
    with torch.no_grad(): net.forward_fused(e, v, 192, None)
for one in (0, 1, 0, 1):
    _lib.set_tuning("mlp_fwd_one_cta", one)
    print(json.dumps({"one_cta_per_sm": one, "fwd_ms": round(timeit(f, 10), 4)}), flush=True)
