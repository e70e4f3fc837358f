"""Host-side data-parallel logic over gloo, world_size 2, on CPU tensors (the CUDA kernels are not involved:
this covers ray sharding, flat-buffer grouping, the summed all-reduce and the 1/world scaling contract)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hn_b200 import dp
        # parameters laid out like HashEmbedder (levels = slices of one buffer) plus a stray tensor
        torch.manual_seed(rank)  # deliberately different: broadcast must fix it
        flat = torch.randn(4, 8, 2)
        levels = [torch.nn.Parameter(flat[i]) for i in range(4)]
        stray = torch.nn.Parameter(torch.randn(5))
        params = levels + [stray]
        dp.broadcast_parameters(params, src=0)
        ref = torch.Generator().manual_seed(0)
        torch.manual_seed(0)
        want_flat = torch.randn(4, 8, 2)
        assert torch.equal(flat, want_flat), "broadcast did not make parameters identical"
        assert all(p.data_ptr() == flat[i].data_ptr() for i, p in enumerate(levels)), "views must stay views"

        # gradients: slices of one flat buffer (what HashEncodeFn.backward returns) + a separate one
        gflat = torch.full((4, 8, 2), float(rank + 1))
        for i, p in enumerate(levels):
            p.grad = gflat[i]
        stray.grad = torch.full((5,), 10.0 * (rank + 1))
        sync = dp.GradSync(params)
        sync.all_reduce()
        sync.wait()
        assert sync.calls_last == 2, f"expected 2 collectives (flat run + stray), got {sync.calls_last}"
        assert sync.bytes_last == (4 * 8 * 2 + 5) * 4
        assert torch.all(gflat == 3.0) and torch.all(stray.grad == 30.0)     # 1 + 2, 10 + 20
        assert sync.grad_scale == 0.5

        # ray sharding covers every ray exactly once
        covered = torch.zeros(11)
        s, e = dp.shard_range(11, rank, world)
        covered[s:e] += 1
        dist.all_reduce(covered)
        assert torch.all(covered == 1)

        # level-bucketed reduction of the flat table gradient: every bucket summed exactly once, nothing else touched
        L_, slab = 6, 10
        tg = torch.arange(L_ * slab, dtype=torch.float32) * (rank + 1)
        red = dp.BucketedTableReducer(L_)
        bk = dp.BucketedTableReducer.buckets(L_, 4)
        assert bk == [(0, 4), (4, 6)]
        red.reduce_levels(tg, *bk[0])
        red.wait()
        want = torch.arange(L_ * slab, dtype=torch.float32)
        assert torch.equal(tg[:4 * slab], 3 * want[:4 * slab]) and torch.equal(tg[4 * slab:], (rank + 1) * want[4 * slab:])
        red.reduce_levels(tg, *bk[1])
        red.wait()
        assert torch.equal(tg, 3 * want) and red.grad_scale == 0.5
        try:
            red.reduce_levels(tg, 2, 9)
            raise AssertionError("out-of-range bucket accepted")
        except ValueError:
            pass

        # inference gather of ragged row blocks
        counts = [dp.shard_range(11, r, world)[1] - dp.shard_range(11, r, world)[0] for r in range(world)]
        local = torch.arange(s, e, dtype=torch.float32)[:, None].repeat(1, 3)
        full = dp.all_gather_rows(local, counts)
        assert torch.equal(full[:, 0], torch.arange(11, dtype=torch.float32))
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_dp_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_shard_range_properties():
    from hn_b200 import dp
    for n in (0, 1, 7, 640000):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_exchange_slices_tile_the_buffer():
    """hn_dp_reduce_update's ownership rule (dp.slice_bounds mirrors csrc/dp_exchange.cu): for every buffer length and
    world size the slices are disjoint, 16-byte aligned and cover [0, n) exactly."""
    from hn_b200 import dp
    for n in (4, 8, 9344, 16 * (1 << 19) * 2, 1000 * 4, 12):
        for world in (1, 2, 3, 4, 7, 8, 64):
            cursor = 0
            for r in range(world):
                b, e = dp.slice_bounds(n, r, world)
                assert b == cursor and b % 4 == 0 and e >= b
                cursor = e
            assert cursor == n
