"""A few full training steps (render_rays 64+128, loss, backward, RAdam) for launch lists / timing."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
import bench

n_rand = int(os.environ.get("N_RAND", 8192))
steps = int(os.environ.get("STEPS", 3))
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
graphed = bool(int(os.environ.get("GRAPH", 0)))
from hn_b200 import _lib, ops
if "SORT_MIN" in os.environ:
    ops.SORT_MIN_POINTS = int(os.environ["SORT_MIN"])
for knob in ("hash_bwd_agg", "hash_fwd_lpg", "hash_bwd_lpg", "hash_agg_max_heads", "mlp_impl"):
    if knob.upper() in os.environ:
        _lib.set_tuning(knob, int(os.environ[knob.upper()]))
from hn_b200 import autograph
if autograph.ENABLED:          # HN_AUTO_GRAPH=1: the eager loop with render_rays replayed as CUDA graphs
    autograph.ensure_stream(dev)
rps, ms = bench.train_step_extra(dev, n_rand, steps=steps, warmup=int(os.environ.get("WARMUP", 2)), graphed=graphed)
print(json.dumps({"n_rand": n_rand, "graphed": graphed, "auto_graph": dict(autograph.stats) if autograph.ENABLED else None, "ms_per_step": round(ms, 3), "rays_per_s": round(rps, 1)}))
