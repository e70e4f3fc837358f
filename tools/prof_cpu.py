"""cProfile of the eager training step (host-side overhead of the drop-in path)."""
import cProfile, pstats, os, sys, io, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
import bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
pr = cProfile.Profile()
orig = bench.time_loop
def timed(fn, steps, warmup, dist=None):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(steps): fn()
    torch.cuda.synchronize()
    pr.disable()
    return 1.0
bench.time_loop = timed
bench.train_step_extra(dev, 1024, steps=30, warmup=5)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats(os.environ.get("SORT", "tottime")).print_stats(45)
print(s.getvalue()[:9000])
