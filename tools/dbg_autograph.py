"""Debug: eager vs HN_AUTO_GRAPH training on the synthetic scene with perturb=1 (RNG inside the captured graphs)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "hashnerf-pytorch_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import test_training_parity as T
from embedding.hash_encoding import HashEmbedder
from embedding.spherical_harmonic import SHEncoder
from hn_b200 import autograph
from loss import total_variation_loss
from models import NeRFSmall
from radam import RAdam
from run_nerf_helpers import render_rays, run_network, img2mse
t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
DEV = "cuda"
perturb = float(os.environ.get("PERTURB", 1.0))
steps, n_rand, log2T = 80, 512, 14
batches = [T.scene_rays(n_rand, 500 + i) for i in range(steps)]
test_rays, test_rgb = T.scene_rays(2048, 4242)
geo = dict(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32, input_ch_views=16)
box = (torch.tensor(T.BBOX[0]), torch.tensor(T.BBOX[1]))

def run():
    torch.manual_seed(3)
    emb = HashEmbedder(box, log2_hashmap_size=log2T, finest_resolution=256).to(DEV)
    c, f, sh = NeRFSmall(**geo).to(DEV), NeRFSmall(**geo).to(DEV), SHEncoder()
    opt = RAdam([{"params": list(c.parameters()) + list(f.parameters()), "weight_decay": 1e-6}, {"params": list(emb.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
    opt.fused_zero_grad = True
    q = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)
    kw = dict(network_fn=c, network_query_fn=q, N_samples=32, embed_fn=emb, retraw=True, perturb=perturb, N_importance=64, network_fine=f, white_bkgd=True, raw_noise_std=0.)
    out = []
    for step, (rays, rgb) in enumerate(batches, start=1):
        ret = render_rays(t(rays).to(DEV), **kw)
        opt.zero_grad()
        tgt = t(rgb).to(DEV)
        loss = img2mse(ret["rgb_map"], tgt) + img2mse(ret["rgb0"], tgt) + 1e-10 * (ret["sparsity_loss"].sum() + ret["sparsity_loss0"].sum())
        tv = sum(total_variation_loss(emb.embeddings[i], 16, 256, i, log2T, n_levels=16) for i in range(16))
        loss = loss + 1e-6 * tv
        loss.backward()
        opt.step()
        if step % 10 == 0:
            with torch.no_grad():
                kw2 = dict(kw, perturb=0.)
                o = render_rays(t(test_rays).to(DEV), **kw2)
            out.append(round(T.psnr(o["rgb_map"].cpu().numpy(), test_rgb), 2))
    return out

print("eager     ", run())
autograph.enable(True); autograph.ensure_stream(torch.device("cuda:0"))
print("autograph ", run(), autograph.stats)
autograph.shutdown()
