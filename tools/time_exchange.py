"""torchrun --nproc-per-node N tools/time_exchange.py : the 64 MiB table-gradient exchange alone, our one-pass kernel
over symmetric memory against the NCCL all-reduce (CUDA events, max over ranks, 30 repetitions after 5 warm-ups)."""
import os, sys, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
from hn_b200 import dp, _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 16 * (1 << 19) * 2


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"world": world, "bytes": n * 4}
x = torch.randn(n, device=dev)
out["nccl_ms"] = round(timed(lambda: dist.all_reduce(x)), 4)
sar = dp.SymmetricAllReduce(n, dev)
sar.tensor.normal_()
for knob in (8, 4, 16, 32):
    _lib.set_tuning("dp_grid_per_sm", knob)
    out[f"ours_ms_grid{knob}"] = round(timed(sar.all_reduce), 4)
_lib.set_tuning("dp_grid_per_sm", 8)
out["multicast"] = sar.multicast
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
