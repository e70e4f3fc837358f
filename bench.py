#!/usr/bin/env python
"""bench.py -- hash-encode fwd+bwd throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 2^24 uniform points in the bbox per GPU, L=16, F=2,
log2_hashmap_size=19, finest_res=512; one step = forward gather + zero-grad + backward scatter (+ one NCCL
all-reduce of the 64 MiB table gradient when N > 1: the data-parallel exchange step the north star names).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hashnerf-pytorch_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

BYTES_PER_SAMPLE_FWD = 12 + 16 * (8 * 2 * 4 + 2 * 4)   # 1164
BYTES_PER_SAMPLE_BWD = 12 + 16 * (2 * 4 + 8 * 2 * 4)   # 1164
BBOX = ((-1.5, -1.5, -1.5), (1.5, 1.5, 1.5))
METRIC = "hash_encode_fwd_bwd_msamples_per_s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def time_loop(fn, steps, warmup, dist=None, finish=None):
    """W untimed + exactly K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks.
    ``finish``: joins whatever the last step left running on a side stream, inside the timed region."""
    for _ in range(warmup):
        fn()
    if finish is not None:
        finish()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's OWN PyTorch code (oracle/_ref, a git-ignored copy made by
# oracle/make_ref.py that travels with the tree; loaded unmodified through oracle/ref_loader.py) on the host
# cores.  Only if that copy is missing does the oracle port stand in (kind "port").
# ------------------------------------------------------------------------------------------------
def _reference(device="cpu"):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    if ref_loader.available():
        return ref_loader.load(device), "reference"
    return None, "port"


def _reference_encoder(ref, log2T, device, seed=0):
    """The reference's HashEmbedder (embedding/hash_encoding.py:14-57) with its own init, on ``device``."""
    torch.manual_seed(seed)
    box = (torch.tensor(BBOX[0], device=device), torch.tensor(BBOX[1], device=device))
    emb = ref.HashEmbedder(box, n_levels=16, n_features_per_level=2, log2_hashmap_size=log2T, base_resolution=16,
                           finest_resolution=512)
    return emb.to(device)


def cpu_hash_encode_fwd_bwd(n_points: int, log2T: int, steps: int, warmup: int):
    """One step = HashEmbedder.forward + autograd backward on ``n_points`` uniform points, all host threads."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref, kind = _reference("cpu")
    g = torch.Generator().manual_seed(0)
    lo, hi = torch.tensor(BBOX[0]), torch.tensor(BBOX[1])
    x = torch.rand(n_points, 3, generator=g) * (hi - lo) + lo
    dy = torch.randn(n_points, 32, generator=g)
    if ref is not None:
        emb = _reference_encoder(ref, log2T, "cpu")

        def step():
            for e in emb.embeddings:
                e.weight.grad = None
            out, _ = emb(x)
            out.backward(dy)
    else:
        import oracle as O
        tables = ((torch.rand(16, 1 << log2T, 2, generator=g) * 2e-4) - 1e-4).requires_grad_(True)
        res = O.level_resolutions()

        def step():
            tables.grad = None
            out, _ = O.hash_encode(x, tables, lo, hi, res, log2T)
            out.backward(dy)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_points / dt / 1e6, dt * 1e3, cores, kind


def calibrated_points(log2T: int, total_steps: int, budget_s: float, cap_log2: int = 19):
    """Largest power-of-two sample of the workload's points whose ``total_steps`` passes fit ``budget_s`` seconds of
    CPU time (measured, not guessed: one pass over 2^15 points is timed first)."""
    _v, ms, _c, _k = cpu_hash_encode_fwd_bwd(1 << 15, log2T, steps=1, warmup=1)
    per_point = ms / 1e3 / (1 << 15)
    k = 15
    while k < cap_log2 and per_point * (1 << (k + 1)) * total_steps <= budget_s:
        k += 1
    return 1 << k


def cpu_pipeline_cfg0(steps: int = 1, warmup: int = 1):
    """BASELINE configs[0]: HashEmbedder + SHEncoder + NeRFSmall + raw2outputs fwd+bwd on CPU, 4096 synthetic rays x
    64 samples, chair hparams, through the reference's own run_network / raw2outputs."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref, kind = _reference("cpu")
    if ref is None:
        return None
    torch.manual_seed(0)
    emb = _reference_encoder(ref, 19, "cpu")
    net = ref.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                        input_ch=32, input_ch_views=16)
    sh = ref.SHEncoder()
    R, S = 4096, 64
    g = torch.Generator().manual_seed(1)
    o = torch.tensor([0., 0., 4.]) + 0.1 * torch.randn(R, 3, generator=g)
    d = -o / o.norm(dim=-1, keepdim=True) + 0.2 * torch.randn(R, 3, generator=g)
    rays = torch.cat([o, d, torch.full((R, 1), 2.), torch.full((R, 1), 6.), d / d.norm(dim=-1, keepdim=True)], -1)
    qfn = lambda i, v, fn: ref.run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh, netchunk=1 << 16)
    target = torch.rand(R, 3, generator=g)

    def step():
        for p_ in list(emb.parameters()) + list(net.parameters()):
            p_.grad = None
        ret = ref.render_rays(rays, net, qfn, S, embed_fn=emb, perturb=1., N_importance=0, white_bkgd=True)
        loss = ((ret["rgb_map"] - target) ** 2).mean() + 1e-10 * ret["sparsity_loss"].sum()
        loss.backward()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"s_per_step": round(dt, 3), "msamples_per_s": round(R * S / dt / 1e6, 4), "rays_per_s": round(R / dt, 1),
            "cores": cores, "kind": kind,
            "workload": "configs[0]: 4096 rays x 64 samples, L=16 F=2 T=2^19 finest 512, coarse pass, "
                        "render_rays fwd + backward through the reference's own modules"}


def reference_gpu_extra(dev, log2T=19):
    """The reference's own PyTorch code on the SAME B200 (SURVEY 2.2 / BASELINE.md 4: the per-kernel bar is the
    reference's chain of ATen kernels on this GPU): hash-encode fwd+bwd at 2^22 points, and the training-loop body
    of run_nerf.py:608-642 (render_rays 64+128, mse + sparsity + 16 TV terms, backward, the reference's RAdam) at
    N_rand 1024 and 8192.  Runs under the CUDA default tensor type exactly as run_nerf.py:725 sets it."""
    out = {}
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref is missing (run oracle/make_ref.py where /root/reference exists)"}
    torch.set_default_tensor_type('torch.cuda.FloatTensor')
    try:
        ref = ref_loader.load("cuda")
        lo, hi = torch.tensor(BBOX[0], device=dev), torch.tensor(BBOX[1], device=dev)
        emb = _reference_encoder(ref, log2T, dev)
        n = 1 << 22
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.rand(n, 3, device=dev, generator=g) * (hi - lo) + lo
        dy = torch.randn(n, 32, device=dev, generator=g)

        def enc_step():
            for e in emb.embeddings:
                e.weight.grad = None
            y, _ = emb(x)
            y.backward(dy)

        ms = time_loop(enc_step, 3, 1) / 3
        out["hash_encode_fwd_bwd_msamples_per_s"] = round(n / ms / 1e3, 2)
        out["hash_encode_points"] = n
        del x, dy
        torch.cuda.empty_cache()

        mk = lambda: ref.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3,
                                   hidden_dim_color=64, input_ch=32, input_ch_views=16).to(dev)
        coarse, fine, sh = mk(), mk(), ref.SHEncoder()
        opt = ref.RAdam([{"params": list(coarse.parameters()) + list(fine.parameters()), "weight_decay": 1e-6},
                         {"params": list(emb.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
        qfn = lambda i, v, fn: ref.run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh, netchunk=1 << 16)
        for n_rand in (1024, 8192):
            o = torch.tensor([0., 0., 4.], device=dev) + 0.1 * torch.randn(n_rand, 3, device=dev, generator=g)
            d = -o / o.norm(dim=-1, keepdim=True) + 0.2 * torch.randn(n_rand, 3, device=dev, generator=g)
            rays = torch.cat([o, d, torch.full((n_rand, 1), 2., device=dev), torch.full((n_rand, 1), 6., device=dev),
                              d / d.norm(dim=-1, keepdim=True)], -1)
            target = torch.rand(n_rand, 3, device=dev, generator=g)

            def train_step():
                ret = ref.render_rays(rays, coarse, qfn, 64, embed_fn=emb, retraw=True, perturb=1., N_importance=128,
                                      network_fine=fine, white_bkgd=True)
                opt.zero_grad()
                loss = ((ret["rgb_map"] - target) ** 2).mean() + ((ret["rgb0"] - target) ** 2).mean() \
                    + 1e-10 * (ret["sparsity_loss"].sum() + ret["sparsity_loss0"].sum())
                tv = sum(ref.total_variation_loss(emb.embeddings[i], 16, 512, i, log2T, n_levels=16) for i in range(16))
                (loss + 1e-6 * tv).backward()
                opt.step()

            ms = time_loop(train_step, 5, 2) / 5
            out[f"train_rays_per_s_nrand{n_rand}"] = round(n_rand / ms * 1e3, 1)
            out[f"train_ms_per_step_nrand{n_rand}"] = round(ms, 2)
        out["what"] = ("the reference's unmodified modules (oracle/_ref) on device='cuda' of this same B200: eager ATen "
                       "kernels, fp32, torch " + torch.__version__)
    except Exception as exc:  # the leg is informative: never lose the bench line over it
        out["error"] = f"{type(exc).__name__}: {exc}"[:300]
    finally:
        torch.set_default_tensor_type('torch.FloatTensor')
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    n = calibrated_points(args.log2T, steps + warmup, budget_s=150.0)
    v, ms, cores, kind = cpu_hash_encode_fwd_bwd(n, args.log2T, steps, warmup)
    what = ("the reference's own HashEmbedder.forward + autograd backward (embedding/hash_encoding.py, unmodified, "
            "torch CPU fp32)" if kind == "reference" else "oracle port of hash_encoding.py fwd + autograd bwd, torch CPU fp32")
    sample = f"{n} of the 2^24 points per step, sized so that {steps}+{warmup} steps fit ~150 s of CPU time; {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n),
        "cpu_baseline": {"value": round(v, 4), "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 4), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_extra:
        line["extra"] = {"cfg0_cpu_pipeline": cpu_pipeline_cfg0()}
    print(json.dumps(line))


def workload_config(args, n_points):
    return {"workload": f"cfg2 hash-encode fwd+bwd: 2^{int(np.log2(args.points))} uniform points in bbox per GPU, L=16, "
                        f"F=2, log2_hashmap_size={args.log2T}, base 16, finest 512",
            "points_per_step_per_gpu": n_points, "log2_hashmap_size": args.log2T,
            "l2": "inputs larger than L2 (x 201 MB + dY 2.1 GB + features 2.1 GB per step); the 64 MiB table stays "
                  "L2-resident across steps, as in steady-state training",
            "timed_path": "value: the launches HashEmbedder.forward + autograd issue for this batch, called directly "
                          "through hn_b200.ops (sort, gather, zero-grad, scatter[, all-reduce]) on device-resident "
                          "inputs; e2e: the same work through the public HashEmbedder API with host inputs",
            "e2e_readback": "a training pipeline keeps features and gradients on the device: the per-step result read "
                            "back is the 16 per-level gradient sums (64 B), not the 2.1 GB feature tensor",
            "parallelism": (f"dp{args.gpus} (points sharded; table gradient all-reduced "
                            + ("once after the scatter, on a side stream under the next step's point sort, joined before "
                               "that step clears the gradient" if getattr(args, "buckets_used", None) == "pipelined"
                               else f"in level buckets {args.buckets_used}, each exchanged on a side stream while the "
                               "next is scattered" if getattr(args, "buckets_used", None)
                               else "in 4 level buckets, each overlapping the next bucket's scatter" if args.bucket_overlap
                               else "once after the scatter") + f"; exchange: {getattr(args, 'exchange_used', args.exchange)})")
            if args.gpus > 1 else "single"}


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------
def train_step_extra(dev, n_rand, steps=8, warmup=3, graphed=False, dist=None, rank=0, world=1, fused_exchange=False,
                     full_graph=False):
    """BASELINE configs[2]/[3] shape: full coarse+fine render + loss + backward + RAdam on synthetic rays.
    With ``dist`` (N > 1): every rank renders its own n_rand rays, gradients are summed over ranks with one flat
    all-reduce per parameter group (hn_b200.dp.GradSync) and 1/world is applied inside the fused RAdam kernel;
    returns whole-job rays/s."""
    from embedding.hash_encoding import HashEmbedder
    from embedding.spherical_harmonic import SHEncoder
    from models import NeRFSmall
    from radam import RAdam
    from run_nerf_helpers import render_rays, run_network, img2mse
    torch.manual_seed(0)
    emb = HashEmbedder((torch.tensor(BBOX[0]), torch.tensor(BBOX[1])), log2_hashmap_size=19).to(dev)
    mk = lambda: NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                           input_ch=32, input_ch_views=16).to(dev)
    coarse, fine, sh = mk(), mk(), SHEncoder()
    opt = RAdam([{"params": list(coarse.parameters()) + list(fine.parameters()), "weight_decay": 1e-6},
                 {"params": list(emb.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
    g = torch.Generator(device=dev).manual_seed(1 + rank)  # identical init (seed 0 above), distinct rays per rank
    o = torch.tensor([0., 0., 4.], device=dev) + 0.1 * torch.randn(n_rand, 3, device=dev, generator=g)
    d = -o / o.norm(dim=-1, keepdim=True) + 0.2 * torch.randn(n_rand, 3, device=dev, generator=g)
    rays = torch.cat([o, d, torch.full((n_rand, 1), 2., device=dev), torch.full((n_rand, 1), 6., device=dev),
                      d / d.norm(dim=-1, keepdim=True)], -1)
    target = torch.rand(n_rand, 3, device=dev, generator=g)
    qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)

    def render_fn(rb):
        return render_rays(rb, coarse, qfn, 64, embed_fn=emb, retraw=True, perturb=1., N_importance=128,
                           network_fine=fine, white_bkgd=True)

    def loss_fn(ret, tgt):
        return img2mse(ret["rgb_map"], tgt) + img2mse(ret["rgb0"], tgt) \
            + 1e-10 * (ret["sparsity_loss"].sum() + ret["sparsity_loss0"].sum())

    if graphed:
        from hn_b200.graph import GraphedTrainStep
        grad_sync = None
        if dist is not None:   # data parallel inside the graph: the all-reduces are captured as graph nodes
            from hn_b200.dp import GradSync
            sync = GradSync(list(emb.parameters()) + list(coarse.parameters()) + list(fine.parameters()))
            opt.grad_scale = sync.grad_scale
            grad_sync = sync.all_reduce_inline
        batcher, graph_loss = None, loss_fn
        use_batcher = full_graph in (True, "batcher")
        use_tv = full_graph in (True, "tv")
        if full_graph:
            # the COMPLETE step in the graph: batch construction (on-device pixel sampling + ray generation + target
            # gather from resident images), render, mse + sparsity + the 16 TV terms, backward, RAdam + zero-grad
            import loss as loss_mod
            from hn_b200.batcher import DeviceRayBatcher
            loss_mod.TV_FAST_DRAWS = True
            H = W = 400
            focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
            K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
            n_img = 8
            images = torch.rand(n_img, H, W, 3, device=dev, generator=g)
            poses = torch.zeros(n_img, 3, 4, device=dev)
            for k in range(n_img):
                a = 2 * np.pi * k / n_img
                poses[k] = torch.tensor([[np.cos(a), 0, np.sin(a), 4 * np.sin(a)], [0, 1, 0, 0],
                                         [-np.sin(a), 0, np.cos(a), 4 * np.cos(a)]], device=dev)
            if use_batcher:
                batcher = DeviceRayBatcher(images, poses, H, W, K, 2., 6., n_rand, dev, precrop_iters=0, seed=rank)

            def graph_loss(ret, tgt):   # all 16 TV terms behind one autograd node (loss.total_variation_sweep)
                return loss_fn(ret, tgt) + 1e-6 * loss_mod.total_variation_sweep(emb).sum()
            if not use_tv:
                graph_loss = loss_fn
        trainer = GraphedTrainStep(n_rand, render_fn, graph_loss, opt, dev, warmup=2, grad_sync=grad_sync, batcher=batcher)

        def step():
            if batcher is not None:
                trainer.step()
            else:
                trainer.step(rays, target)
        for _ in range(4):  # eager warm-up + capture happen outside the timed region
            step()
    else:
        from loss import total_variation_loss
        sync, fx = None, None
        if dist is not None and fused_exchange:
            from hn_b200.dp import FusedExchange
            fx = FusedExchange(opt, [emb, coarse, fine])
        elif dist is not None:
            from hn_b200.dp import GradSync
            sync = GradSync(list(emb.parameters()) + list(coarse.parameters()) + list(fine.parameters()))

        def step():  # the statements of the reference's loop body, run_nerf.py:608-642
            ret = render_fn(rays)
            opt.zero_grad()
            loss = loss_fn(ret, target)
            tv = sum(total_variation_loss(emb.embeddings[i], emb.base_resolution, emb.finest_resolution, i,
                                          emb.log2_hashmap_size, n_levels=emb.n_levels) for i in range(emb.n_levels))
            loss = loss + 1e-6 * tv
            loss.backward()
            if fx is not None:      # collective + RAdam + zero_grad in one pass over peer memory
                fx.step()
                return
            if sync is not None:
                sync.all_reduce()
                sync.wait()
                opt.grad_scale = sync.grad_scale
            opt.step()

    ms = time_loop(step, steps, warmup, dist) / steps
    return world * n_rand / ms * 1e3, ms


def inference_frame_extra(dev, H=800, W=800, dist=None):
    """BASELINE configs[4] shape: one 800x800 frame = 640 000 rays x (64 + 128) samples, perturb 0, no_grad."""
    from embedding.hash_encoding import HashEmbedder
    from embedding.spherical_harmonic import SHEncoder
    from models import NeRFSmall
    from run_nerf_helpers import render, run_network
    torch.manual_seed(0)
    emb = HashEmbedder((torch.tensor(BBOX[0]), torch.tensor(BBOX[1])), log2_hashmap_size=19).to(dev)
    mk = lambda: NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                           input_ch=32, input_ch_views=16).to(dev)
    coarse, fine, sh = mk(), mk(), SHEncoder()
    qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = torch.tensor([[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 4.0]], device=dev)

    kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=coarse, network_fine=fine,
              network_query_fn=qfn, N_samples=64, N_importance=128, embed_fn=emb, perturb=0., raw_noise_std=0.,
              white_bkgd=True)

    if dist is None:
        def frame():
            with torch.no_grad():
                render(H, W, K, chunk=1024 * 32, c2w=c2w, **kw)
    else:
        # SURVEY 8e inference partition: a contiguous range of the frame's rays per rank, one all-gather at the end
        from hn_b200.dp import render_image_sharded

        def render_fn(o, d):
            rgb, depth, acc, _extras = render(H, W, K, chunk=1024 * 32, rays=(o, d), **kw)
            return rgb, depth, acc

        def frame():
            with torch.no_grad():
                render_image_sharded(H, W, K, c2w, render_fn)

    ms = time_loop(frame, 3, 1, dist) / 3
    return ms


def fern_frame_extra(dev):
    """The LLFF half of BASELINE configs[4]: fern.txt shapes -- a 378 x 504 frame (factor 8), N_samples 64 +
    N_importance 64, forward-facing NDC rays (near 0, far 1, hn_ndc_rays), raw_noise_std 0 at test time, no white
    background (configs/fern.txt:9-14; run_nerf.py:246-247)."""
    from embedding.hash_encoding import HashEmbedder
    from embedding.spherical_harmonic import SHEncoder
    from models import NeRFSmall
    from run_nerf_helpers import render, run_network
    torch.manual_seed(0)
    H, W, focal = 378, 504, 407.6
    box = (torch.tensor([-1.6, -1.6, -1.1]), torch.tensor([1.6, 1.6, 1.1]))   # the bbox.py rule for NDC scenes: +-(1 + margin)
    emb = HashEmbedder(box, log2_hashmap_size=19).to(dev)
    mk = lambda: NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                           input_ch=32, input_ch_views=16).to(dev)
    coarse, fine, sh = mk(), mk(), SHEncoder()
    qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = torch.tensor([[1, 0, 0, 0.05], [0, 1, 0, -0.02], [0, 0, 1, 0.3]], device=dev)
    kw = dict(ndc=True, near=0., far=1., use_viewdirs=True, network_fn=coarse, network_fine=fine, network_query_fn=qfn,
              N_samples=64, N_importance=64, embed_fn=emb, perturb=0., raw_noise_std=0., white_bkgd=False)

    def frame():
        with torch.no_grad():
            render(H, W, K, chunk=1024 * 32, c2w=c2w, **kw)
    return time_loop(frame, 3, 1) / 3


def dp_self_check(dist, dev, world, bwd, dflat, tables, exchange_step=None):
    """N > 1, once before timing: (1) the all-reduced table gradient's checksum equals the sum of the per-rank
    checksums gathered separately; (2) after one fused RAdam step on the reduced gradient (1/world folded in) the
    parameters are bit-identical on every rank; (3) the scatter + exchange exactly as the timed step issues them
    (``exchange_step``: our symmetric-memory kernels, bucketed or not) leave the same reduced gradient in place as
    the NCCL all-reduce of a copy.  Works on copies: the timed state is untouched."""
    from radam import RAdam
    dflat.zero_()
    bwd()
    local = dflat.double().sum().reshape(1)
    sums = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(sums, local)
    expect = float(sum(t.item() for t in sums))
    red = dflat.clone()
    dist.all_reduce(red)
    got = float(red.double().sum().item())
    scale = float(red.double().abs().sum().item()) + 1e-30
    ok_sum = abs(got - expect) <= 1e-6 * scale
    p = torch.nn.Parameter(tables.detach().reshape(-1).clone())
    p.grad = red.reshape(-1)
    opt = RAdam([{"params": [p], "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
    opt.grad_scale = 1.0 / world
    opt.step()
    h = p.detach().view(torch.int32).long().sum().reshape(1)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    ok_par = all(int(t.item()) == int(hs[0].item()) for t in hs)
    ok_ex = True
    if exchange_step is not None:
        dflat.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier()
        exchange_step()
        torch.cuda.synchronize(dev)
        ok_ex = float((dflat - red).abs().max().item()) <= 1e-5 * float(red.abs().max().item())
    dflat.zero_()
    return "ok" if (ok_sum and ok_par and ok_ex) else \
        f"FAILED (checksum {ok_sum}, parameters identical {ok_par}, own exchange equals NCCL {ok_ex})"


def train_shape_roofline(dev, emb, peak, steps=10):
    """The kernels the TRAINING path launches (caller-ordered ray samples, no sort, warp-aggregated scatter) at the
    data-parallel shape: 8192 rays x (64 + 192) samples = 2,097,152 points along rays; same 1164 B/sample formula."""
    from hn_b200 import ops
    R, S = 8192, 256
    g = torch.Generator(device=dev).manual_seed(3)
    o = torch.tensor([0., 0., 4.], device=dev) + 0.1 * torch.randn(R, 3, device=dev, generator=g)
    d = -o / o.norm(dim=-1, keepdim=True) + 0.2 * torch.randn(R, 3, device=dev, generator=g)
    z = torch.sort(2. + 4. * torch.rand(R, S, device=dev, generator=g), -1).values
    pts = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3).contiguous()
    n = pts.shape[0]
    dy = torch.randn(n, 32, device=dev, generator=g)
    tables = emb.flat_tables()
    box, res = emb._geometry(dev)
    dflat = torch.zeros(tables.numel(), device=dev)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)   # > L2: the points / dY / features are evicted

    def fwd():
        flush.zero_()
        ops.hash_encode_forward(pts, tables, box, res, 16, 2, emb.log2_hashmap_size, want_keep=True)

    def bwd():
        flush.zero_()
        ops.hash_encode_backward(pts, dy, box, res, 16, 2, emb.log2_hashmap_size, dflat, ordered=True)

    def only_flush():
        flush.zero_()

    t_flush = time_loop(only_flush, steps, 2) / steps
    t_f = time_loop(fwd, steps, 2) / steps - t_flush
    t_b = time_loop(bwd, steps, 2) / steps - t_flush
    gb_f, gb_b = n * BYTES_PER_SAMPLE_FWD / t_f / 1e6, n * BYTES_PER_SAMPLE_BWD / t_b / 1e6
    return {"points": n, "shape": "8192 rays x (64 + 192) samples, ray order (hn_hash_encode_fwd / _bwd_ordered)",
            "l2": "192 MiB written between launches (its time subtracted)",
            "fwd": {"ms": round(t_f, 4), "gbs": round(gb_f, 1), "frac": round(gb_f / peak, 4)},
            "bwd": {"ms": round(t_b, 4), "gbs": round(gb_b, 1), "frac": round(gb_b / peak, 4)}, "peak": peak, "unit": "GB/s"}


def run_ours(args):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    from hn_b200 import _lib, ops
    from embedding.hash_encoding import HashEmbedder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    n, log2T, L, F = args.points, args.log2T, 16, 2
    gen = torch.Generator(device=dev).manual_seed(rank)
    lo, hi = torch.tensor(BBOX[0], device=dev), torch.tensor(BBOX[1], device=dev)
    x = torch.rand(n, 3, device=dev, generator=gen) * (hi - lo) + lo
    dy = torch.randn(n, L * F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    torch.manual_seed(0)  # identical tables on every rank
    emb = HashEmbedder((lo.cpu(), hi.cpu()), log2_hashmap_size=log2T).to(dev)
    tables = emb.flat_tables()
    box, res = emb._geometry(dev)
    # N > 1: the table gradient lives in symmetric memory and is summed by our own one-pass kernel (multimem
    # reduce through the NVSwitch + multicast of the sum, csrc/dp_exchange.cu); --exchange nccl keeps the library call
    sar, exchange = None, "none"
    if dist is not None:
        exchange = "nccl"
        if args.exchange in ("auto", "nvls"):
            try:
                from hn_b200.dp import SymmetricAllReduce
                sar = SymmetricAllReduce(tables.numel(), dev)
                exchange = "hn_dp_reduce_update over symmetric memory (" + ("NVLS multimem" if sar.multicast else "peer loads/stores") + ")"
            except Exception as exc:
                if args.exchange == "nvls":
                    raise
                exchange = f"nccl (symmetric memory unavailable: {type(exc).__name__})"
    dflat = sar.tensor if sar is not None else torch.zeros(tables.numel(), device=dev)
    out_holder = {}

    grid_res = ops.sort_grid_res(n)

    def sort():
        out_holder["xs4"] = ops.hash_sort_points(x, box, grid_res)

    def fwd():
        out_holder["y"] = ops.hash_encode_forward_sorted(out_holder["xs4"], tables, box, res, L, F, log2T,
                                                         want_keep=True)[0]

    def bwd():
        ops.hash_encode_backward_sorted(out_holder["xs4"], dy, box, res, L, F, log2T, dflat)

    reducer, buckets, overlap, pipelined = None, [], None, False
    if dist is not None and sar is not None and args.buckets != "none":
        from hn_b200.dp import OverlappedTableReducer
        overlap = OverlappedTableReducer(sar, L)
        if args.buckets == "pipelined":
            pipelined = True
        else:
            buckets = [tuple(int(v) for v in part.split("-")) for part in args.buckets.split(",")]
            assert buckets[0][0] == 0 and buckets[-1][1] == L and all(a[1] == b[0] for a, b in zip(buckets, buckets[1:])), \
                "--buckets must tile [0, L)"
        args.buckets_used = args.buckets
    elif dist is not None and sar is None and args.bucket_overlap:
        from hn_b200.dp import BucketedTableReducer
        reducer = BucketedTableReducer(L)
        buckets = BucketedTableReducer.buckets(L, 4)

    def step():
        # what HashEmbedder.forward + autograd backward launch for this many points: counting sort by grid
        # cell, sorted gather, zero-grad, warp-aggregated scatter (+ the DP all-reduce)
        if pipelined:
            # the exchange of step s runs through the switch on a side stream while step s+1 sorts ITS points (the
            # sort depends on the batch only); it is joined before the gradient buffer is cleared and before the
            # gather reads the tables an optimizer would have updated from it
            sort()
            overlap.wait()
            dflat.zero_()
            fwd()
            bwd()
            overlap.reduce_levels(0, L)
            return
        dflat.zero_()
        sort()
        fwd()
        if dist is None:
            bwd()
        elif overlap is not None:
            # level buckets over symmetric memory: bucket b is exchanged through the switch (side stream) while
            # bucket b+1 is scattered; only the last bucket's exchange is exposed
            for b, e in buckets:
                ops.hash_encode_backward_sorted(out_holder["xs4"], dy, box, res, L, F, log2T, dflat, levels=(b, e))
                overlap.reduce_levels(b, e)
            overlap.wait()
        elif reducer is None:
            bwd()
            if sar is not None:
                sar.all_reduce()
            else:
                dist.all_reduce(dflat)
        else:
            # level buckets: the all-reduce of bucket b runs on a side stream while bucket b+1 is scattered
            for b, e in buckets:
                ops.hash_encode_backward_sorted(out_holder["xs4"], dy, box, res, L, F, log2T, dflat, levels=(b, e))
                reducer.reduce_levels(dflat, b, e)
            reducer.wait()

    args.exchange_used = exchange
    dp_check = None
    if dist is not None:
        out_holder["xs4"] = ops.hash_sort_points(x, box, grid_res)
        def exchange_step():   # the scatter + exchange part of step(), on an already zeroed dflat
            if pipelined:
                bwd()
                overlap.reduce_levels(0, L)
                overlap.wait()
            elif overlap is not None:
                for b, e in buckets:
                    ops.hash_encode_backward_sorted(out_holder["xs4"], dy, box, res, L, F, log2T, dflat, levels=(b, e))
                    overlap.reduce_levels(b, e)
                overlap.wait()
            else:
                bwd()
                sar.all_reduce()
        dp_check = dp_self_check(dist, dev, world, bwd, dflat, tables, exchange_step if sar is not None else None)

    # ---- headline: device-resident inputs
    clocks = ClockSampler(local)
    launches0 = _lib.launches
    if rank == 0:
        clocks.start()
    total_ms = time_loop(step, args.steps, args.warmup, dist, finish=overlap.wait if overlap is not None else None)
    clock_report = clocks.stop() if rank == 0 else None
    gpu_launches = (_lib.launches - launches0) * args.steps // (args.steps + args.warmup)
    ms_per_step = total_ms / args.steps
    value = world * n / ms_per_step / 1e3  # Msamples/s, whole job

    # ---- per-kernel timing for the roofline (same stream, CUDA events)
    sort_ms = time_loop(sort, args.steps, 2) / args.steps
    fwd_ms = time_loop(fwd, args.steps, 2) / args.steps
    bwd_ms = time_loop(bwd, args.steps, 2) / args.steps
    peak, peak_src = peaks()
    dom = "hash_bwd_kernel" if bwd_ms >= fwd_ms else "hash_fwd_kernel"
    dom_ms = max(bwd_ms, fwd_ms)
    dom_bytes = n * (BYTES_PER_SAMPLE_BWD if dom == "hash_bwd_kernel" else BYTES_PER_SAMPLE_FWD)
    achieved = dom_bytes / dom_ms / 1e6  # GB/s
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": round(dom_ms, 4),
                "fwd": {"ms": round(fwd_ms, 4), "gbs": round(n * BYTES_PER_SAMPLE_FWD / fwd_ms / 1e6, 1)},
                "bwd": {"ms": round(bwd_ms, 4), "gbs": round(n * BYTES_PER_SAMPLE_BWD / bwd_ms / 1e6, 1)},
                "sort": {"ms": round(sort_ms, 4), "grid_res": grid_res,
                         "note": "counting sort of the points by grid cell (histogram, scan, partition with one global cursor per bin, per-bin local sort), counted in the step"},
                "step_frac_of_hbm": round(n * (BYTES_PER_SAMPLE_FWD + BYTES_PER_SAMPLE_BWD)
                                          / (sort_ms + fwd_ms + bwd_ms) / 1e6 / peak, 4),
                "traffic_source": "profiles/traffic.json (dram__bytes_read + write per launch from the ncu --set full "
                                  "capture of this build; not re-measured in this run)"}
    roofline_train = train_shape_roofline(dev, emb, peak, steps=max(5, args.steps // 2)) if rank == 0 else None

    # ---- end to end through the public API with host inputs
    # The user-level pattern for host-resident points (a data loader with one step of prefetch): every step
    # uploads ITS input points from pinned host memory into one of two device buffers on a copy stream; the
    # upload of step s+1 is issued when step s starts computing, so it overlaps step s's kernels.  Every timed
    # step therefore contains one full H2D upload (201 MB), the encode + scatter through HashEmbedder.forward /
    # autograd, and the D2H read of the step's result; the piecewise variant (--e2e-chunks > 1) instead splits
    # one step's points into geometric pieces and overlaps piece c+1's upload with piece c's compute.
    x_host = x.cpu().pin_memory()
    result_host = torch.empty(L, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    n_chunks = max(1, args.e2e_chunks)

    if sar is not None:
        # the public API accumulates into the encoder's gradient sink: let that be the symmetric buffer
        if overlap is not None:
            overlap.wait()
        torch.cuda.synchronize(dev)
        sar.tensor.zero_()
        emb.grad_sink().adopt(sar.tensor)

    def finish_step(main):
        flat = emb.grad_sink().flat
        if dist is not None:                                     # the same exchange as the device-resident step
            if sar is not None:
                sar.all_reduce()
            else:
                dist.all_reduce(flat)
        g = flat.view(L, -1).sum(dim=1)
        result_host.copy_(g, non_blocking=True)                  # D2H of the step's result (per-level grad sums)
        main.synchronize()

    if n_chunks == 1:
        x_bufs = [torch.empty_like(x), torch.empty_like(x)]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"cur": 0}

        def upload(b):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])              # the kernels that read this buffer are done
                x_bufs[b].copy_(x_host, non_blocking=True)       # H2D of one step's input points
                copied[b].record(copy_stream)

        upload(0)                                                # prologue: the first step's input

        def e2e_step():
            main = torch.cuda.current_stream()
            b = state["cur"]
            for e in emb.embeddings:
                e.weight.grad = None
            main.wait_event(copied[b])
            upload(b ^ 1)                                        # next step's input, overlapping this step
            feats, _keep = emb(x_bufs[b])                        # HashEmbedder.forward (public API)
            feats.backward(dy)                                   # autograd -> scatter kernel
            consumed[b].record(main)
            state["cur"] = b ^ 1
            finish_step(main)

        api = ("HashEmbedder.forward + autograd backward; each step's points are uploaded from pinned host memory "
               "into one of two device buffers on a copy stream, issued one step ahead (prefetch depth 1) so the "
               "upload overlaps the previous step's kernels; upstream gradient dY resident (stands for the "
               "downstream MLP); per-level gradient sums read back every step")
    else:
        x_dev = torch.empty_like(x)
        edges = [0] + [n * ((1 << (i + 1)) - 1) // ((1 << n_chunks) - 1) for i in range(n_chunks)]
        bounds = [(edges[i], edges[i + 1]) for i in range(n_chunks)]
        copied = [torch.cuda.Event() for _ in range(n_chunks)]
        consumed = [torch.cuda.Event() for _ in range(n_chunks)]

        def e2e_step():
            main = torch.cuda.current_stream()
            for e in emb.embeddings:
                e.weight.grad = None
            for c, (a, b) in enumerate(bounds):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[c])
                    x_dev[a:b].copy_(x_host[a:b], non_blocking=True)
                    copied[c].record(copy_stream)
            for c, (a, b) in enumerate(bounds):
                main.wait_event(copied[c])
                feats, _keep = emb(x_dev[a:b])
                feats.backward(dy[a:b])
                consumed[c].record(main)
            finish_step(main)

        api = ("HashEmbedder.forward + autograd backward per piece (sizes 1:2:4:...), piece c+1 uploaded from pinned "
               "host memory on a copy stream while piece c computes; dY resident; per-level gradient sums read back")

    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = time_loop(e2e_step, e2e_steps, 2, dist) / e2e_steps
    torch.cuda.synchronize()
    if dist is not None:
        api += ("; N > 1: the table gradient is exchanged every step (" + exchange + ") before the sums are taken -- "
                "in line, nothing overlapped: the host waits for each step's result")
    e2e = {"value": round(world * n / e2e_ms / 1e3, 2), "unit": "Msamples/s", "h2d_bytes_per_step": n * 12,
           "d2h_bytes_per_step": L * 4, "ms_per_step": round(e2e_ms, 3), "chunks": n_chunks, "api": api}

    # ---- N > 1: the data-parallel training step (every rank takes part)
    extra = {}
    if world > 1 and not args.no_extra:
        x_host = x_dev = x_bufs = None
        out_holder.clear()
        torch.cuda.empty_cache()
        rps, ms = train_step_extra(dev, 8192, steps=10, warmup=3, dist=dist, rank=rank, world=world)
        extra["train_rays_per_s_nrand8192_per_rank_dp"] = round(rps, 1)
        extra["train_ms_per_step_nrand8192_per_rank_dp"] = round(ms, 3)
        extra["train_step"] = ("data parallel: 8192 rays per rank, render_rays 64+128, mse+sparsity+16 TV terms, "
                               "backward, flat gradient all-reduce (GradSync), RAdam with 1/world folded in; eager")
        if sar is not None:
            try:
                rps, ms = train_step_extra(dev, 8192, steps=10, warmup=3, dist=dist, rank=rank, world=world,
                                           fused_exchange=True)
                extra["train_rays_per_s_nrand8192_per_rank_dp_fused_exchange"] = round(rps, 1)
                extra["train_ms_per_step_nrand8192_per_rank_dp_fused_exchange"] = round(ms, 3)
                extra["train_step_fused_exchange"] = ("same step, but all-reduce + RAdam + zero_grad are ONE pass over "
                                                      "peer memory (hn_b200.dp.FusedExchange: multimem reduce, sharded "
                                                      "moments, parameter multicast)")
            except Exception as exc:
                extra["train_step_fused_exchange"] = f"failed: {type(exc).__name__}: {exc}"[:200]
        if args.dp_graph:
            rps, ms = train_step_extra(dev, 8192, steps=20, warmup=3, graphed=True, dist=dist, rank=rank, world=world)
            extra["train_rays_per_s_nrand8192_per_rank_dp_cuda_graph"] = round(rps, 1)
            extra["train_ms_per_step_nrand8192_per_rank_dp_cuda_graph"] = round(ms, 3)
        ms = inference_frame_extra(dev, dist=dist)
        extra["inference_800x800_ms_per_frame_rays_sharded"] = round(ms, 2)
        extra["inference_800x800_mrays_per_s_rays_sharded"] = round(0.64 / ms * 1e3, 2)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- extras (rank 0 only, N=1 only): other table sizes, training step
    if world == 1 and not args.no_extra:
        x_host = x_dev = x_bufs = None  # release the e2e buffers before the extra workloads
        for t_log2 in (14, 22):
            e2 = HashEmbedder((lo.cpu(), hi.cpu()), log2_hashmap_size=t_log2).to(dev)
            tb, (bx, rs) = e2.flat_tables(), e2._geometry(dev)
            dg = torch.zeros(tb.numel(), device=dev)

            def s2():
                dg.zero_()
                xs = ops.hash_sort_points(x, bx, grid_res)
                ops.hash_encode_forward_sorted(xs, tb, bx, rs, L, F, t_log2, want_keep=True)
                ops.hash_encode_backward_sorted(xs, dy, bx, rs, L, F, t_log2, dg)
            m = time_loop(s2, max(3, args.steps // 2), 2) / max(3, args.steps // 2)
            extra[f"msamples_per_s_T{t_log2}"] = round(n / m / 1e3, 1)
            del e2, tb, dg
        def plain():
            dflat.zero_()
            ops.hash_encode_forward(x, tables, box, res, L, F, log2T, want_keep=True)
            ops.hash_encode_backward(x, dy, box, res, L, F, log2T, dflat)
        m = time_loop(plain, max(3, args.steps // 2), 2) / max(3, args.steps // 2)
        extra["msamples_per_s_T19_unsorted_path"] = round(n / m / 1e3, 1)
        del dy
        out_holder.clear()
        torch.cuda.empty_cache()
        for n_rand in (1024, 8192):
            rps, ms = train_step_extra(dev, n_rand)
            extra[f"train_rays_per_s_nrand{n_rand}"] = round(rps, 1)
            extra[f"train_ms_per_step_nrand{n_rand}"] = round(ms, 3)
            rps, ms = train_step_extra(dev, n_rand, steps=20, graphed=True)
            extra[f"train_rays_per_s_nrand{n_rand}_cuda_graph"] = round(rps, 1)
            extra[f"train_ms_per_step_nrand{n_rand}_cuda_graph"] = round(ms, 3)
        ms = inference_frame_extra(dev)
        extra["inference_800x800_ms_per_frame"] = round(ms, 2)
        extra["inference_800x800_mrays_per_s"] = round(0.64 / ms * 1e3, 2)
        ms = fern_frame_extra(dev)
        extra["inference_fern_378x504_ndc_64p64_ms_per_frame"] = round(ms, 2)
        extra["inference_fern_378x504_ndc_64p64_mrays_per_s"] = round(378 * 504 / ms / 1e3, 2)
        torch.cuda.empty_cache()
        extra["reference_gpu"] = reference_gpu_extra(dev, log2T)
        for n_rand in (1024, 8192):   # last: a failed capture must not be able to disturb the other legs
            try:
                rps, ms = train_step_extra(dev, n_rand, steps=20, graphed=True, full_graph=True)
                extra[f"train_rays_per_s_nrand{n_rand}_cuda_graph_full"] = round(rps, 1)
                extra[f"train_ms_per_step_nrand{n_rand}_cuda_graph_full"] = round(ms, 3)
            except Exception as exc:
                extra[f"train_nrand{n_rand}_cuda_graph_full"] = f"failed: {type(exc).__name__}: {exc}"[:200]
                break
            finally:
                import loss as _loss_mod
                _loss_mod.TV_FAST_DRAWS = False
        # the eager loop again, with render_rays replayed as CUDA graphs UNDER the drop-in API (HN_AUTO_GRAPH=1: what an
        # unmodified run_nerf.py gets with that variable set); switches the process to a side stream, hence last
        from hn_b200 import autograph
        try:
            autograph.enable(True)
            autograph.ensure_stream(dev)
            for n_rand in (1024, 8192):
                rps, ms = train_step_extra(dev, n_rand, steps=20, warmup=6)
                extra[f"train_rays_per_s_nrand{n_rand}_auto_graph"] = round(rps, 1)
                extra[f"train_ms_per_step_nrand{n_rand}_auto_graph"] = round(ms, 3)
            extra["auto_graph_stats"] = dict(autograph.stats)
        except Exception as exc:
            extra["train_auto_graph"] = f"failed: {type(exc).__name__}: {exc}"[:200]
        finally:
            autograph.shutdown()
        extra["train_step"] = ("render_rays 64+128 samples/ray, perturb=1, white_bkgd, mse+sparsity, backward, RAdam"
                               "; eager = the drop-in API driven like run_nerf.py:608-642 incl. the 16 TV-loss terms, cuda_graph "
                               "= hn_b200.graph.GraphedTrainStep replaying render+loss+backward+RAdam(+zero-grad) on given rays (no TV); "
                               "cuda_graph_full = the same graph with the on-device ray batcher (pixel sampling, ray "
                               "generation, target gather from 8 resident 400x400 images) and the 16 TV terms inside; auto_graph = the eager "
                               "loop statements unchanged, render_rays replayed as a forward and a backward CUDA graph behind "
                               "one autograd node (hn_b200.autograph, HN_AUTO_GRAPH=1)")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = calibrated_points(log2T, 4, budget_s=20.0, cap_log2=20)
        v, ms, cores, kind = cpu_hash_encode_fwd_bwd(n_cpu, log2T, steps=3, warmup=1)
        cpu = {"value": round(v, 4), "unit": "Msamples/s", "cores": cores, "kind": kind,
               "sample": f"{n_cpu} of the 2^24 points, 3 timed passes after 1 warm-up ({ms:.0f} ms/pass), "
                         + ("the reference's own HashEmbedder.forward + autograd backward (oracle/_ref, unmodified)"
                            if kind == "reference" else "oracle port of the reference's PyTorch encoder fwd + autograd bwd")}

    line = {"metric": METRIC, "value": round(value, 2), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "roofline": roofline, "roofline_train": roofline_train, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(gpu_launches), "clocks": clock_report, "extra": extra}
    if dp_check is not None:
        line["dp_check"] = dp_check
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2T", type=int, default=19)
    ap.add_argument("--points", type=int, default=1 << 24)
    ap.add_argument("--e2e-chunks", type=int, default=1,
                    help="1 = upload each step's points one step ahead (double buffer); >1 = split one step into "
                         "geometric pieces and overlap piece c+1's upload with piece c's compute")
    ap.add_argument("--dp-graph", action="store_true",
                    help="N > 1: also time the data-parallel training step as one CUDA graph with the NCCL all-reduces "
                         "captured inside (opt-in: keeps the default run free of collective capture)")
    ap.add_argument("--bucket-overlap", action="store_true",
                    help="N > 1, --exchange nccl: all-reduce the table gradient in 4 level buckets overlapped with the "
                         "scatter instead of once after it (measured slower with NCCL: splitting the scatter costs "
                         "more than the overlap hides)")
    ap.add_argument("--buckets", default="pipelined",
                    help="N > 1, symmetric-memory exchange.  'pipelined' (default): the exchange of step s runs on a "
                         "side stream under the point sort of step s+1 and is joined before that step clears the "
                         "gradient / gathers; 'none': one exchange after the scatter, nothing overlapped; 'b-e,b-e': "
                         "level buckets, each exchanged while the next is scattered (measured slower: the split "
                         "scatter re-reads dY)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nvls", "nccl"],
                    help="N > 1: how gradients are exchanged -- our one-pass kernel over symmetric memory (auto: with "
                         "fall-back to NCCL if symmetric memory cannot be set up) or the NCCL library all-reduce")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
