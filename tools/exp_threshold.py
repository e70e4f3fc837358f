"""Where does sorting start to pay?  fwd+bwd through the public API, random points, coherent on/off."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from embedding.hash_encoding import HashEmbedder
from sweep_hash import timeit
dev = torch.device("cuda:0")
emb = HashEmbedder((torch.tensor([-1.5] * 3), torch.tensor([1.5] * 3)), log2_hashmap_size=19).to(dev)
for logn in (14, 16, 17, 18, 19, 20, 22):
    n = 1 << logn
    x = torch.rand(n, 3, device=dev) * 3 - 1.5
    dy = torch.randn(n, 32, device=dev)
    row = {"log2N": logn}
    for coh in (False, True):
        emb.coherent = coh
        def fb():
            out, _ = emb(x); out.backward(dy)
        row["sorted_ms" if coh else "plain_ms"] = round(timeit(fb, 10), 4)
    print(json.dumps(row), flush=True)
