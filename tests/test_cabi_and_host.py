"""CPU-side checks: the C-ABI library loads and exports exactly what include/hashnerf_b200.h declares, the
ctypes table mirrors it, and the host-side mirrors of the reference modules behave (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "hashnerf_b200.h")).read()
    return sorted(set(re.findall(r"HN_API\s+[\w\s\*]+?\b(hn_\w+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("hn_hash_encode_fwd", "hn_hash_encode_bwd", "hn_hash_sort_points", "hn_hash_encode_fwd_sorted",
                 "hn_hash_encode_bwd_sorted", "hn_spatial_hash", "hn_voxel_vertices", "hn_sh_encode", "hn_mlp_fwd",
                 "hn_mlp_bwd", "hn_composite_fwd", "hn_composite_bwd", "hn_sample_pdf", "hn_sort_concat_rows",
                 "hn_coarse_z", "hn_ray_points", "hn_radam_step", "hn_last_error_string", "hn_abi_version"):
        assert must in names, must


def test_library_exports_every_declared_symbol():
    from hn_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        from hn_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.hn_abi_version() == 2
    # argument validation happens before any CUDA call, so it is testable without a GPU
    lib.hn_last_error_string.restype = ctypes.c_char_p
    rc = lib.hn_sh_encode(None, ctypes.c_int64(4), 9, None, None)
    assert rc == -22 and b"degree" in lib.hn_last_error_string()
    rc = lib.hn_hash_encode_fwd(None, None, None, None, ctypes.c_int64(8), 99, 2, 19, None, None, None)
    assert rc == -22
    assert lib.hn_sh_encode(None, ctypes.c_int64(0), 4, None, None) == 0  # empty input is a no-op


def test_ctypes_table_matches_header():
    from hn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    text = open(os.path.join(ROOT, "include", "hashnerf_b200.h")).read()
    for name, (_res, args) in _lib.SIGNATURES.items():
        decl = re.search(r"HN_API[^;(]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", text, re.S).group(1)
        n_args = 0 if decl.strip() in ("", "void") else decl.count(",") + 1
        assert n_args == len(args), f"{name}: header has {n_args} parameters, ctypes table {len(args)}"


def test_tuning_keys_documented_in_header():
    """Every key hn_set_tuning accepts is listed in the header's comment, and nothing else is."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "hashnerf-pytorch_b200", "csrc", "hash_encode.cu")).read()
    accepted = set(re.findall(r'strcmp\(key, "([a-z0-9_]+)"\)', src))
    header = open(os.path.join(root, "include", "hashnerf_b200.h")).read()
    block = header[header.index("Launch-shape knobs"):header.index("HN_API int hn_set_tuning")]
    documented = set(re.findall(r'"([a-z0-9_]+)"', block))
    assert accepted and accepted == documented, (sorted(accepted - documented), sorted(documented - accepted))


def test_sort_grid_rule_and_level_buckets():
    from hn_b200 import ops
    from hn_b200.dp import BucketedTableReducer
    for n in (1, 1000, 1 << 19, 1 << 20, 1 << 22, 1 << 24, 1 << 30):
        g = ops.sort_grid_res(n)
        assert 16 <= g <= 256 and (g & (g - 1)) == 0
    assert ops.sort_grid_res(1 << 24) == 256 and ops.sort_grid_res(1 << 21) == 128
    assert BucketedTableReducer.buckets(16, 4) == [(0, 4), (4, 8), (8, 12), (12, 16)]
    assert BucketedTableReducer.buckets(5, 2) == [(0, 2), (2, 4), (4, 5)]


def test_hash_embedder_host_side():
    from embedding.hash_encoding import HashEmbedder, HASH_PRIMES
    import embedding.hash_encoding as he
    assert HASH_PRIMES[:3] == [1, 2654435761, 805459861]
    assert he.BOX_OFFSETS.shape == (1, 8, 3) and he.BOX_OFFSETS[0, 5].tolist() == [1, 0, 1]
    torch.manual_seed(3)
    emb = HashEmbedder((torch.zeros(3), torch.ones(3)), log2_hashmap_size=8)
    assert emb.out_dim == 32 and emb.n_levels == 16 and emb.log2_hashmap_size == 8
    assert int(emb.base_resolution) == 16 and int(emb.finest_resolution) == 512
    keys = list(emb.state_dict().keys())
    assert keys == [f"embeddings.{i}.weight" for i in range(16)]
    assert all(e.weight.shape == (256, 2) for e in emb.embeddings)
    assert float(emb.embeddings[0].weight.detach().abs().max()) <= 1e-4
    flat = emb.flat_tables()
    assert flat.shape == (16, 256, 2) and flat.data_ptr() == emb.embeddings[0].weight.data_ptr()
    # identical construction order as the reference => identical tables for a seeded RNG
    torch.manual_seed(3)
    ref = [torch.nn.Embedding(256, 2) for _ in range(16)]
    for r in ref:
        torch.nn.init.uniform_(r.weight, a=-0.0001, b=0.0001)
    assert all(torch.equal(r.weight, e.weight) for r, e in zip(ref, emb.embeddings))
    # a state_dict round trip and a dtype/device-style _apply keep the single flat buffer
    other = HashEmbedder((torch.zeros(3), torch.ones(3)), log2_hashmap_size=8)
    other.load_state_dict(emb.state_dict())
    assert torch.equal(other.flat_tables(), flat)
    other.float()
    from hn_b200 import ops
    assert ops._consecutive([e.weight for e in other.embeddings])
    assert emb.level_resolutions().tolist() == [16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        emb(torch.rand(4, 3))
    # TV loss applies embeddings[i] to an index tensor (loss.py:30): plain nn.Embedding semantics
    assert emb.embeddings[2](torch.tensor([[0, 5]])).shape == (1, 2, 2)


def test_hash_embedder_copies_and_pickles():
    """deepcopy / pickle (EMA copies, torch.save(module)): the copy owns its level modules -- the back reference that
    total_variation_loss uses must lead to the COPY's tables and gradient buffer, and nothing unpicklable is stored."""
    import copy
    import io
    from embedding.hash_encoding import HashEmbedder, level_owner
    emb = HashEmbedder((torch.zeros(3), torch.ones(3)), log2_hashmap_size=8)
    assert level_owner(emb.embeddings[3]) == (emb, 3)
    assert level_owner(torch.nn.Embedding(4, 2)) == (None, None)
    twin = copy.deepcopy(emb)
    assert level_owner(twin.embeddings[3]) == (twin, 3) and level_owner(emb.embeddings[3]) == (emb, 3)
    assert twin.flat_tables().data_ptr() != emb.flat_tables().data_ptr()
    assert torch.equal(twin.flat_tables(), emb.flat_tables())
    buf = io.BytesIO()
    torch.save(emb, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert level_owner(back.embeddings[15]) == (back, 15)
    assert torch.equal(back.flat_tables(), emb.flat_tables()) and back.grad_sink() is not emb.grad_sink()
    # a level module moved into another container no longer claims its old encoder's slot
    stolen = emb.embeddings[2]
    emb.embeddings[2] = torch.nn.Embedding(256, 2)
    assert level_owner(stolen) == (None, None)


def test_nerf_small_and_sh_host_side():
    from embedding.spherical_harmonic import SHEncoder
    from models import NeRF, NeRFSmall
    net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                    input_ch=32, input_ch_views=16)
    assert list(net.state_dict().keys()) == ["sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight",
                                             "color_net.1.weight", "color_net.2.weight"]
    assert [tuple(p.shape) for p in net.parameters()] == [(64, 32), (16, 64), (64, 31), (64, 64), (3, 64)]
    assert net.flat_weights().numel() == 9344
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(4, 48))
    with pytest.raises(NotImplementedError):
        NeRF()
    sh = SHEncoder()
    assert sh.out_dim == 16 and sh.degree == 4
    with pytest.raises(AssertionError):
        SHEncoder(degree=6)


def test_radam_interface_matches_reference_layout():
    from radam import RAdam
    p = torch.nn.Parameter(torch.zeros(3))
    opt = RAdam([{"params": [p], "weight_decay": 1e-6}], lr=0.01, betas=(0.9, 0.99))
    g = opt.param_groups[0]
    assert set(g) >= {"lr", "betas", "eps", "weight_decay", "buffer", "params"}
    assert len(g["buffer"]) == 10
    for bad in (dict(lr=-1), dict(eps=-1), dict(betas=(1.0, 0.9)), dict(betas=(0.9, 1.0))):
        with pytest.raises(ValueError):
            RAdam([p], **bad)
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
    assert opt._rectification(1, 0.9, 0.99)[0] == 0 and opt._rectification(6, 0.9, 0.99)[0] == 1


def test_helpers_import_surface():
    import run_nerf_helpers as h
    for name in ("create_nerf", "render", "render_path", "render_rays", "run_network", "batchify", "raw2outputs",
                 "sample_pdf", "get_embedder", "img2mse", "mse2psnr", "to8b", "device", "HashEmbedder", "SHEncoder",
                 "NeRFSmall", "NeRF", "NeRFGradient", "RAdam", "get_rays", "get_rays_np", "get_ndc_rays"):
        assert hasattr(h, name), name
    import inspect
    assert list(inspect.signature(h.render_rays).parameters) == [
        "ray_batch", "network_fn", "network_query_fn", "N_samples", "embed_fn", "retraw", "lindisp", "perturb",
        "N_importance", "network_fine", "white_bkgd", "raw_noise_std", "verbose", "pytest"]
    assert list(inspect.signature(h.raw2outputs).parameters) == ["raw", "z_vals", "rays_d", "raw_noise_std",
                                                                 "white_bkgd", "pytest"]
    assert list(inspect.signature(h.sample_pdf).parameters) == ["bins", "weights", "N_samples", "det", "pytest"]
    from embedding.embedder import Embedder, get_embedder  # noqa: F401  (reference import line :19)
    from loss import sigma_sparsity_loss, total_variation_loss  # noqa: F401
    assert float(h.mse2psnr(torch.tensor(0.01))) == pytest.approx(20.0, abs=1e-4)


def test_autograph_is_opt_in_and_inert_without_cuda():
    """hn_b200.autograph only acts when HN_AUTO_GRAPH=1 (or enable()) AND its side stream exists: by default, and on a
    machine without CUDA, render_rays runs the eager statements and nothing switches streams."""
    import inspect
    import run_nerf_helpers as h
    from hn_b200 import autograph
    assert autograph.ENABLED is (os.environ.get("HN_AUTO_GRAPH", "0") == "1")
    if not autograph.ENABLED:
        assert autograph.ensure_stream("cuda") is None
        assert autograph.render_rays(h._render_rays_eager, torch.zeros(4, 11), {}) is None
    # the eager implementation keeps the public function's parameters (the wrapper forwards them by name)
    pub = list(inspect.signature(h.render_rays).parameters)
    assert list(inspect.signature(h._render_rays_eager).parameters) == pub
    autograph.reset()
    assert set(autograph.stats) >= {"captures", "captures_bwd", "replays", "eager", "failed"}


def test_sort_workspace_covers_both_layouts():
    """hn_hash_sort_workspace_bytes is host arithmetic: it must cover the two-level layout (histogram, first slots,
    one 32-byte sector per bin cursor, N float4 records) and the single-pass layout (one counter per cell, key and
    rank per point), whichever the call ends up using."""
    from hn_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.hn_hash_sort_workspace_bytes.restype = ctypes.c_int64
    lib.hn_hash_sort_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
    bins, stride = 16384, 8
    for n, g in [(0, 1), (1, 8), (70_001, 32), (1 << 19, 64), (1 << 24, 256), (1 << 24, 300), (1000, 1024)]:
        got = lib.hn_hash_sort_workspace_bytes(n, g)
        two_level = (2 * (bins + 4) + bins * stride) * 4 + n * 16
        single = (g ** 3 + (g ** 3 + 2047) // 2048 + 2 * n) * 4
        assert got >= two_level and got >= single, (n, g, got, two_level, single)
        assert got % 4 == 0
    assert lib.hn_hash_sort_workspace_bytes(-1, 8) == -1 and lib.hn_hash_sort_workspace_bytes(8, 0) == -1
    assert lib.hn_hash_sort_workspace_bytes(8, 1025) == -1
