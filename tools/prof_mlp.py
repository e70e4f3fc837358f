"""One forward + backward of the tcgen05 MLP at training size, for ncu captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))
from models import NeRFSmall
dev = torch.device("cuda:0")
net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16).to(dev)
N = 8192 * 192
e = (torch.randn(N, 32, device=dev) * 0.3).requires_grad_(True)
v = torch.randn(8192, 16, device=dev)
dO = torch.randn(N, 4, device=dev)
for _ in range(2):
    out = net.forward_fused(e, v, 192, None)
    out.backward(dO)
torch.cuda.synchronize()
print("ok")
