"""Dense tensor-core peaks of this box, measured the way MEASURED_PEAKS.json measures bf16 (cuBLAS through
torch.matmul, 8192^3, best of 10 with CUDA events): TF32 (the forward MLP's MMA kind) and bf16 (the fused backward's).
The MLP kernels' tensor-pipe percentages in profiles/ are quoted against these."""
import json
import torch

dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0), "how": "torch.matmul 8192^3 (2*N^3 flop), best of 10, CUDA events"}
n = 8192
for name, dtype, tf32 in (("tf32_tflops", torch.float32, True), ("bf16_tflops", torch.bfloat16, False),
                          ("fp32_ffma_tflops", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dtype)
    b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(2):
        a @ b
    best = 1e9
    for _ in range(10 if name != "fp32_ffma_tflops" else 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name] = round(2 * n ** 3 / best / 1e9, 1)
torch.backends.cuda.matmul.allow_tf32 = False
print(json.dumps(out))
