import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
import bench
from hn_b200 import ops, _lib
if "SORT_MIN" in os.environ:
    ops.SORT_MIN_POINTS = int(os.environ["SORT_MIN"])
for knob in ("hash_fwd_lpg", "mlp_impl", "hash_level_major"):
    if knob.upper() in os.environ:
        _lib.set_tuning(knob, int(os.environ[knob.upper()]))
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
print(json.dumps({"sort_min": ops.SORT_MIN_POINTS, "ms_per_frame": round(bench.inference_frame_extra(dev), 2)}))
