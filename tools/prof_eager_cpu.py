"""Host-side profile (cProfile) of the eager N_rand = 1024 training step: where the Python / launch time goes."""
import cProfile, io, os, pstats, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
import bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
n_rand = int(os.environ.get("N_RAND", 1024))
pr = cProfile.Profile()
orig = bench.time_loop

def profiled(fn, steps, warmup, dist=None, finish=None):
    for _ in range(warmup + 5):
        fn()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(steps):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    return orig(fn, steps, 1, dist, finish)

bench.time_loop = profiled
from hn_b200 import autograph
if autograph.ENABLED:
    autograph.ensure_stream(dev)
rps, ms = bench.train_step_extra(dev, n_rand, steps=30, warmup=3)
print(f"eager step {ms:.3f} ms")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats(os.environ.get("SORT", "tottime")).print_stats(40)
print(s.getvalue()[:6000])
