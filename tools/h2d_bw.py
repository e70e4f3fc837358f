import torch, time
x = torch.empty(1<<24, 3).pin_memory(); d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): d.copy_(x, non_blocking=True)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print("H2D 201MB ms", ms, "GB/s", x.numel() * 4 / ms / 1e6)
