import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib, ops
from models import NeRFSmall
from sweep_hash import timeit
dev = torch.device("cuda:0")
net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16).to(dev)
for rays, S in ((8192, 192), (8192, 64), (1024, 192)):
    N = rays * S
    e = (torch.randn(N, 32, device=dev) * 0.3).requires_grad_(True); v = torch.randn(rays, 16, device=dev); dO = torch.randn(N, 4, device=dev)
    def fb():
        out = net.forward_fused(e, v, S, None); out.backward(dO)
    def f():
        with torch.no_grad(): net.forward_fused(e, v, S, None)
    res = {"N": N, "fwd_ms": round(timeit(f, 5), 3)}
    grads = {}
    for nb in (2, 1):
        _lib.set_tuning("mlp_dw_nbuf", nb)
        res[f"fwd_bwd_ms_nbuf{nb}"] = round(timeit(fb, 5), 3)
        net.zero_grad(); e.grad = None; fb(); grads[nb] = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
    res["grad_rel_diff"] = float((grads[1] - grads[2]).abs().max() / grads[2].abs().max())
    print(json.dumps(res), flush=True)
