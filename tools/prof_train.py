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
rps, ms = bench.train_step_extra(dev, n_rand, steps=steps, warmup=int(os.environ.get("WARMUP", 2)), graphed=graphed)
print(json.dumps({"n_rand": n_rand, "graphed": graphed, "ms_per_step": round(ms, 3), "rays_per_s": round(rps, 1)}))
