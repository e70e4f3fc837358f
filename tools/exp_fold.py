"""MLP forward: folded middle layers (4 rounds) vs layer by layer (5 rounds): time and agreement."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib
from models import NeRFSmall
from sweep_hash import timeit
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16).to(dev)
for rays, S in ((8192, 192), (8192, 64), (1024, 192), (1024, 64)):
    N = rays * S
    e = torch.randn(N, 32, device=dev) * 0.3; v = torch.randn(rays, 16, device=dev)
    def f():
        with torch.no_grad(): return net.forward_fused(e, v, S, None)
    res = {"N": N}
    outs = {}
    for fold in (0, 1, 0, 1):
        _lib.set_tuning("mlp_fwd_fold", fold)
        outs[fold] = f().clone()
        res[f"fold{fold}_ms"] = round(timeit(f, 10), 4)
    # fp64 reference
    with torch.no_grad():
        W = [p.double() for p in net.parameters()]
        x = e.double(); sh = v.double().repeat_interleave(S, 0)
        h1 = torch.relu(x @ W[0].T); h2 = h1 @ W[1].T
        c = torch.cat([sh, h2[:, 1:]], -1)
        h = torch.relu(c @ W[2].T); h = torch.relu(h @ W[3].T); rgb = h @ W[4].T
        ref = torch.cat([rgb, h2[:, :1]], -1)
    for fold in (0, 1):
        res[f"fold{fold}_max_abs_err"] = float((outs[fold].double() - ref).abs().max())
    res["ref_max_abs"] = float(ref.abs().max())
    print(json.dumps(res), flush=True)
