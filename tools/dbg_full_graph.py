import os, sys, traceback, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/hashnerf-pytorch_b200')
import bench
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
for mode in ("tv", True):
    try:
        print(mode, bench.train_step_extra(dev, 1024, steps=5, graphed=True, full_graph=mode), flush=True)
    except Exception as e:
        traceback.print_exc()
        break
