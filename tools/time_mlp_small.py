"""Warm, back-to-back launch times of hn_mlp_fwd / hn_mlp_bwd through the C-ABI at small and large point counts
(the per-launch fixed cost of the tcgen05 kernels: image staging, TMEM allocation, pipeline fill, dW flush)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
from hn_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
w = (torch.randn(ops.MLP_PARAMS, device=dev, generator=g) * 0.1)
s = torch.cuda.current_stream().cuda_stream
for N, ppv in ((256, 64), (37888, 64), (65536, 64), (196608, 192), (524288, 64), (1572864, 192)):
    enc = torch.randn(N, 32, device=dev, generator=g) * 0.3
    views = torch.randn(N // ppv, 16, device=dev, generator=g)
    dout = torch.randn(N, 4, device=dev, generator=g)
    out = torch.empty(N, 4, device=dev); gates = torch.empty(N, 6, dtype=torch.int32, device=dev)
    d_enc = torch.empty(N, 32, device=dev); dflat = torch.zeros(ops.MLP_PARAMS, device=dev)
    ws = torch.empty(max(1, lib.hn_mlp_bwd_workspace_bytes(N) // 4), device=dev)

    def fwd():
        _lib.call("hn_mlp_fwd", enc.data_ptr(), 32, views.data_ptr(), 16, ppv, w.data_ptr(), None, N, out.data_ptr(),
                  gates.data_ptr(), s)

    def bwd():
        _lib.call("hn_mlp_bwd", enc.data_ptr(), 32, views.data_ptr(), 16, ppv, w.data_ptr(), None, gates.data_ptr(),
                  dout.data_ptr(), N, d_enc.data_ptr(), dflat.data_ptr(), ws.data_ptr(), s)

    res = {"N": N}
    for name, fn, abl in (("fwd", fwd, 0), ("bwd", bwd, 0), ("bwd_noflush", bwd, 1), ("bwd_noprep", bwd, 2),
                          ("bwd_neither", bwd, 3)):
        _lib.set_tuning("mlp_dw_ablate", abl)
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        reps = 50
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        res[name + "_us"] = round(a.elapsed_time(b) / reps * 1e3, 1)
    _lib.set_tuning("mlp_dw_ablate", 0)
    print(json.dumps(res), flush=True)
