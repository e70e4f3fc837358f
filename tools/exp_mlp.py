"""MLP implementations: accuracy against an fp64 evaluation and timing (FFMA vs tcgen05)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hn_b200 import _lib, ops
import cases, oracle as O
from models import NeRFSmall
from sweep_hash import timeit
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
sig, col = cases.mlp_weights(5)
net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16)
with torch.no_grad():
    for lin, w in zip(list(net.sigma_net) + list(net.color_net), sig + col):
        lin.weight.copy_(T(w))
net.to(dev)
rs = np.random.RandomState(0)
n, S = 192 * 300 + 77, 1
enc = (rs.randn(n, 32) * 0.3).astype(np.float32); views = rs.randn(n, 16).astype(np.float32)
x64 = torch.cat([T(enc), T(views)], -1).double()
ref = O.nerf_small(x64, [T(w).double() for w in sig], [T(w).double() for w in col]).numpy()
ref32 = O.nerf_small(x64.float(), [T(w) for w in sig], [T(w) for w in col]).numpy()
print(json.dumps({"cpu_fp32_vs_fp64_max_abs": float(np.abs(ref32 - ref).max()), "out_absmax": float(np.abs(ref).max())}))
for impl in (0, 1):
    _lib.set_tuning("mlp_impl", impl)
    with torch.no_grad():
        out = net.forward_fused(T(enc).to(dev), T(views).to(dev), 1, None).cpu().numpy()
    err = np.abs(out - ref)
    rel = err / np.maximum(np.abs(ref), 1e-3)
    print(json.dumps({"impl": impl, "max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()), "max_rel_err(floor 1e-3)": float(rel.max()),
                      "frac_within_rtol1e-5_atol1e-6": float(np.isclose(out, ref32, rtol=1e-5, atol=1e-6).mean())}), flush=True)
# timing at training size
N = 8192 * 192
e = torch.randn(N, 32, device=dev) * 0.3; v = torch.randn(8192, 16, device=dev)
for impl in (0, 1):
    _lib.set_tuning("mlp_impl", impl)
    with torch.no_grad():
        t = timeit(lambda: net.forward_fused(e, v, 192, None), 10)
    print(json.dumps({"impl": impl, "fwd_ms_1.57Mpts": round(t, 4), "Mpts_per_s": round(N / t / 1e3, 1), "tflops_algorithmic": round(N * 18688 / t / 1e9, 2)}), flush=True)

# ---- backward accuracy vs an fp64 autograd evaluation, and timing
n2, ppv = 192 * 40, 192
enc2 = (rs.randn(n2, 32) * 0.3).astype(np.float32); views2 = rs.randn(n2 // ppv, 16).astype(np.float32)
dout = rs.randn(n2, 4).astype(np.float32); keepm = rs.rand(n2) > 0.05
ws64 = [T(w).double().requires_grad_(True) for w in sig + col]
e64 = T(enc2).double().requires_grad_(True)
full = torch.cat([e64, T(views2).double().repeat_interleave(ppv, 0)], -1)
o = O.nerf_small(full, ws64[:2], ws64[2:])
o = torch.cat([o[:, :3], torch.where(T(keepm), o[:, 3], torch.zeros((), dtype=torch.double))[:, None]], -1)
(o * T(dout).double()).sum().backward()
for impl in (0, 1):
    _lib.set_tuning("mlp_impl", impl)
    for lin in list(net.sigma_net) + list(net.color_net):
        lin.weight.grad = None
    eg = T(enc2).to(dev).requires_grad_(True)
    out = net.forward_fused(eg, T(views2).to(dev), ppv, T(keepm).to(dev))
    (out * T(dout).to(dev)).sum().backward()
    res = {"impl": impl, "d_enc_err_rel_to_max": float((eg.grad.cpu().double() - e64.grad).abs().max() / e64.grad.abs().max())}
    for i, (lin, w64) in enumerate(zip(list(net.sigma_net) + list(net.color_net), ws64)):
        res[f"dW{i}_err_rel_to_max"] = float((lin.weight.grad.cpu().double() - w64.grad).abs().max() / w64.grad.abs().max())
    print(json.dumps(res), flush=True)
dO = torch.randn(N, 4, device=dev)
for impl in (0, 1):
    _lib.set_tuning("mlp_impl", impl)
    eg = e.clone().requires_grad_(True)
    def fb():
        out = net.forward_fused(eg, v, 192, None)
        out.backward(dO)
    t = timeit(fb, 5)
    print(json.dumps({"impl": impl, "fwd_bwd_ms_1.57Mpts": round(t, 4), "tflops_algorithmic": round(N * 54016 / t / 1e9, 2)}), flush=True)
