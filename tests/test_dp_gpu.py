"""Data parallelism on real GPUs (needs >= 2): two NCCL ranks, each rendering its shard of the rays, must end
up -- after GradSync.all_reduce -- with the gradient a single process computes on the whole batch, and after
RAdam(grad_scale = 1/world) with identical parameters on both ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dev):
    import cases
    from embedding.hash_encoding import HashEmbedder
    from embedding.spherical_harmonic import SHEncoder
    from models import NeRFSmall
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    tables = cases.synth_tables(16, 12, 2) * np.float32(3000.0)
    emb = HashEmbedder((torch.tensor(cases.BBOX_UNIT[0]), torch.tensor(cases.BBOX_UNIT[1])), log2_hashmap_size=12)
    with torch.no_grad():
        for l in range(16):
            emb.embeddings[l].weight.copy_(T(tables[l]))
    emb.to(dev)
    nets = []
    for seed in (1, 2):
        sig, col = cases.mlp_weights(seed)
        net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                        input_ch=32, input_ch_views=16)
        with torch.no_grad():
            for lin, w in zip(list(net.sigma_net) + list(net.color_net), sig + col):
                lin.weight.copy_(T(w))
        nets.append(net.to(dev))
    return emb, nets, SHEncoder()


def _loss_and_backward(emb, nets, sh, rays):
    from run_nerf_helpers import render_rays, run_network
    qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)
    ret = render_rays(rays, nets[0], qfn, 16, embed_fn=emb, retraw=True, perturb=0., N_importance=16,
                      network_fine=nets[1], white_bkgd=True)
    loss = ret["rgb_map"].square().sum() + ret["rgb0"].square().sum() + 1e-3 * ret["sparsity_loss"].sum()
    loss.backward()   # a SUM over rays: per-rank gradients add up to the full-batch gradient
    return loss


def _graphed_dp_check(rank, world, dev, rays, s, e):
    """The data-parallel exchange inside a CUDA graph (opt-in, HN_TEST_DP_GRAPH=1: capturing NCCL collectives makes
    this test take about two minutes, most of it in communicator teardown).  Passed bit-identically on 2 x B200 when
    committed; it is the test that exposed the stale-scalar race fixed in radam.RAdam.graph_prepare."""
    import torch.distributed as dist
    from hn_b200 import dp
    from radam import RAdam
    from hn_b200.graph import GraphedTrainStep
    from run_nerf_helpers import render_rays as _rr, run_network as _rn, img2mse
    emb3, nets3, sh3 = _build(dev)
    params3 = list(emb3.parameters()) + [p for n in nets3 for p in n.parameters()]
    dp.broadcast_parameters(params3)
    start3 = torch.cat([p.detach().reshape(-1) for p in params3]).clone()
    opt3 = RAdam([{"params": [p for n in nets3 for p in n.parameters()], "weight_decay": 1e-6},
                  {"params": list(emb3.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
    sync3 = dp.GradSync(params3)
    opt3.grad_scale = sync3.grad_scale
    q3 = lambda i, v, fn: _rn(i, v, fn, embed_fn=emb3, embeddirs_fn=sh3)
    render_fn = lambda rb: _rr(rb, nets3[0], q3, 16, embed_fn=emb3, retraw=True, perturb=1., N_importance=16,
                               network_fine=nets3[1], white_bkgd=True)
    loss_fn = lambda ret, tgt: img2mse(ret["rgb_map"], tgt) + img2mse(ret["rgb0"], tgt)
    n_local = e - s
    trainer = GraphedTrainStep(n_local, render_fn, loss_fn, opt3, dev, warmup=2, grad_sync=sync3.all_reduce_inline)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    for _ in range(8):
        pick = torch.randperm(rays.shape[0], device=dev, generator=gen)[:n_local]
        loss3 = trainer.step(rays[pick].contiguous(), torch.rand(n_local, 3, device=dev, generator=gen))
    assert trainer.graph is not None and bool(torch.isfinite(loss3))
    flat3 = torch.cat([p.detach().reshape(-1) for p in params3])
    other3 = [torch.empty_like(flat3) for _ in range(world)]
    dist.all_gather(other3, flat3)
    same3 = all(torch.equal(o, other3[0]) for o in other3)
    assert same3, "ranks diverged under the graphed data-parallel step"
    assert not torch.equal(flat3, start3), "graphed steps did not train"


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import cases
        from hn_b200 import dp
        from radam import RAdam
        rays = torch.from_numpy(cases.rays(96, 7)).to(dev)
        emb, nets, sh = _build(dev)
        params = list(emb.parameters()) + [p for n in nets for p in n.parameters()]
        dp.broadcast_parameters(params)
        s, e = dp.shard_range(rays.shape[0], rank, world)
        _loss_and_backward(emb, nets, sh, rays[s:e].contiguous())
        sync = dp.GradSync(params)
        sync.all_reduce()
        sync.wait()
        assert sync.calls_last == 3, f"expected 3 flat all-reduces (tables, coarse MLP, fine MLP), got {sync.calls_last}"
        got = torch.cat([p.grad.reshape(-1) for p in params]).clone()
        # single-process reference on the whole batch, same device
        emb2, nets2, sh2 = _build(dev)
        _loss_and_backward(emb2, nets2, sh2, rays)
        want = torch.cat([p.grad.reshape(-1) for p in list(emb2.parameters()) + [p for n in nets2 for p in n.parameters()]])
        err = (got - want).abs().max().item() / want.abs().max().item()
        assert err < 1e-4, f"sharded + all-reduced gradient differs from the full-batch gradient: {err}"
        # averaged update, identical on every rank
        opt = RAdam([{"params": [p for n in nets for p in n.parameters()], "weight_decay": 1e-6},
                     {"params": list(emb.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99),
                    degenerated_to_sgd=True)
        opt.grad_scale = sync.grad_scale
        opt.step()
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        assert all(torch.equal(o, other[0]) for o in other), "ranks diverged after the optimizer step"
        if os.environ.get("HN_TEST_DP_GRAPH") == "1":
            _graphed_dp_check(rank, world, dev, rays, s, e)
        # inference partition: rays of a frame sharded over ranks + all-gather == the frame rendered by one process
        from run_nerf_helpers import render, run_network
        H, W = 20, 24
        focal = 0.5 * W / np.tan(0.5 * 0.69)
        K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
        c2w = torch.tensor([[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 2.5]], device=dev)
        qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb2, embeddirs_fn=sh2)
        kw = dict(ndc=False, near=1., far=4., use_viewdirs=True, network_fn=nets2[0], network_fine=nets2[1],
                  network_query_fn=qfn, N_samples=16, N_importance=16, embed_fn=emb2, perturb=0., raw_noise_std=0.,
                  white_bkgd=True)

        def render_fn(o, d):
            rgb, depth, acc, _ = render(H, W, K, chunk=128, rays=(o, d), **kw)
            return rgb, depth, acc
        with torch.no_grad():
            rgb_s, depth_s, acc_s = dp.render_image_sharded(H, W, K, c2w, render_fn)
            rgb_1, depth_1, acc_1, _ = render(H, W, K, chunk=4096, c2w=c2w, **kw)
        assert rgb_s.shape == (H, W, 3) and depth_s.shape == (H, W)
        for a, b, what in ((rgb_s, rgb_1, "rgb"), (depth_s, depth_1, "depth"), (acc_s, acc_1, "acc")):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=True), f"sharded frame differs: {what}"
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_gradient_equivalence_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def _fused_worker(rank, world, port, q):
    """FusedExchange (one pass over peer memory: reduce + RAdam on the owned slice + parameter multicast + gradient
    clear) against GradSync.all_reduce + RAdam.step on the same batches; and SymmetricAllReduce against NCCL."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import cases
        from hn_b200 import dp
        from radam import RAdam
        rays = torch.from_numpy(cases.rays(96, 7)).to(dev)
        s, e = dp.shard_range(rays.shape[0], rank, world)

        def make():
            emb, nets, sh = _build(dev)
            params = list(emb.parameters()) + [p for n in nets for p in n.parameters()]
            dp.broadcast_parameters(params)
            opt = RAdam([{"params": [p for n in nets for p in n.parameters()], "weight_decay": 1e-6},
                         {"params": list(emb.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
            return emb, nets, sh, params, opt

        # A: library all-reduce + the fused RAdam kernel
        emb_a, nets_a, sh_a, params_a, opt_a = make()
        sync = dp.GradSync(params_a)
        opt_a.grad_scale = sync.grad_scale
        # B: the fused exchange
        emb_b, nets_b, sh_b, params_b, opt_b = make()
        fx = dp.FusedExchange(opt_b, [emb_b] + nets_b)
        steps = 7   # crosses RAdam's rectification switch (N_sma >= 5 from step 6 on)
        for k in range(steps):
            lr = 0.01 * (0.1 ** (k / 5.0))
            for g in opt_a.param_groups + opt_b.param_groups:
                g["lr"] = lr
            batch = rays[s:e].contiguous() * (1.0 + 0.01 * k)
            opt_a.zero_grad()
            _loss_and_backward(emb_a, nets_a, sh_a, batch)
            sync.all_reduce()
            sync.wait()
            opt_a.step()
            opt_b.zero_grad()                      # what a run_nerf-style loop does; the buffers are already clear
            _loss_and_backward(emb_b, nets_b, sh_b, batch)
            fx.step()
            torch.cuda.synchronize()
            assert float(fx.grads_sym.tensor.abs().max()) == 0.0, "gradients must be cleared by the exchange"
        flat_a = torch.cat([p.detach().reshape(-1) for p in params_a])
        flat_b = torch.cat([p.detach().reshape(-1) for p in params_b])
        err = float((flat_a - flat_b).abs().max()) / float(flat_a.abs().max())
        assert err < 2e-6, f"fused exchange drifted from all-reduce + RAdam: {err}"   # summation order differs only
        other = [torch.empty_like(flat_b) for _ in range(world)]
        dist.all_gather(other, flat_b)
        assert all(torch.equal(o, other[0]) for o in other), "ranks diverged under the fused exchange"
        # sharded moments reassemble to what the unsharded optimizer holds
        m, v = fx.gather_moments()
        m_a = torch.cat([opt_a.state[p]["exp_avg"].reshape(-1) for p in list(emb_a.parameters())])
        n_tab = m_a.numel()
        assert float((m[:n_tab] - m_a).abs().max()) <= 2e-6 * float(m_a.abs().max()) + 1e-12
        # plain all-reduce through the same kernel
        sar = dp.SymmetricAllReduce(1 << 20, dev)
        gen = torch.Generator(device=dev).manual_seed(5 + rank)
        x = torch.randn(1 << 20, device=dev, generator=gen)
        want = x.clone()
        dist.all_reduce(want)
        for _ in range(3):
            sar.tensor.copy_(x)
            sar.all_reduce()
        torch.cuda.synchronize()
        assert float((sar.tensor - want).abs().max()) <= 1e-6 * float(want.abs().max())
        q.put((rank, f"ok multicast={fx.multicast}"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()[-2500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_exchange_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    print(results)
    assert all(msg.startswith("ok") for _r, msg in results), results


def _overlap_worker(rank, world, port, q):
    """The table-gradient exchange in level buckets on a side stream (hn_b200.dp.OverlappedTableReducer) must leave
    the same sum on every rank as NCCL's all-reduce of the per-rank gradients, also over several steps (the barrier
    epochs count calls)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hashnerf-pytorch_b200"))
        from hn_b200 import ops
        from hn_b200.dp import OverlappedTableReducer, SymmetricAllReduce
        L, F, log2T, n = 16, 2, 14, 1 << 19
        box = torch.tensor([-1.5] * 3 + [1.5] * 3, device=dev)
        res = torch.tensor([16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512.0], device=dev)
        sar = SymmetricAllReduce(L * (1 << log2T) * F, dev)
        ov = OverlappedTableReducer(sar, L)
        dflat = sar.tensor
        msg = "ok"
        for step, buckets in enumerate(([(0, 12), (12, 16)], [(0, 4), (4, 8), (8, 16)], [(0, 16)])):
            g = torch.Generator(device=dev).manual_seed(100 * step + rank)
            x = torch.rand(n, 3, device=dev, generator=g) * 3 - 1.5
            dy = torch.randn(n, L * F, device=dev, generator=g)
            xs4 = ops.hash_sort_points(x, box, 64)
            own = torch.zeros_like(dflat)
            ops.hash_encode_backward_sorted(xs4, dy, box, res, L, F, log2T, own)
            want = own.clone()
            dist.all_reduce(want)
            dflat.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier()
            for b, e in buckets:
                ops.hash_encode_backward_sorted(xs4, dy, box, res, L, F, log2T, dflat, levels=(b, e))
                ov.reduce_levels(b, e)
            ov.wait()
            torch.cuda.synchronize(dev)
            err = (dflat - want).abs().max().item() / want.abs().max().item()
            if not err <= 1e-5:
                msg = f"step {step} buckets {buckets}: relative difference {err:.3e}"
                break
        with pytest.raises(ValueError):
            sar.all_reduce(2, 10)
        q.put((rank, msg))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, f"{type(exc).__name__}: {exc}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_overlapped_table_exchange_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    print(results)
    assert all(msg.startswith("ok") for _r, msg in results), results
