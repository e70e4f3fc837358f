"""Where does the end-to-end step lose time against the resident step?  Chunked compute with and without uploads."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from embedding.hash_encoding import HashEmbedder
from sweep_hash import timeit
dev = torch.device("cuda:0")
n = 1 << 24
box = (torch.tensor([-1.5] * 3), torch.tensor([1.5] * 3))
emb = HashEmbedder(box).to(dev)


gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(n, 3, device=dev, generator=gen) * 3 - 1.5
dy = torch.randn(n, 32, device=dev, generator=gen)
x_host = x.cpu().pin_memory(); x_dev = torch.empty_like(x)
copy_stream = torch.cuda.Stream(device=dev)
def run(n_chunks, upload):
    bounds = [(i * n // n_chunks, (i + 1) * n // n_chunks) for i in range(n_chunks)]
    copied = [torch.cuda.Event() for _ in range(n_chunks)]; consumed = [torch.cuda.Event() for _ in range(n_chunks)]
    def step():
        main = torch.cuda.current_stream()
        for e in emb.embeddings: e.weight.grad = None
        if upload:
            for c, (a, b) in enumerate(bounds):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[c]); x_dev[a:b].copy_(x_host[a:b], non_blocking=True); copied[c].record(copy_stream)
        for c, (a, b) in enumerate(bounds):
            if upload: main.wait_event(copied[c])
            feats, _ = emb((x_dev if upload else x)[a:b]); feats.backward(dy[a:b]); consumed[c].record(main)
        main.synchronize()
    return timeit(step, 5)
for nc in (1, 2, 4, 8):
    print(json.dumps({"chunks": nc, "resident_ms": round(run(nc, False), 3), "upload_ms": round(run(nc, True), 3)}), flush=True)
