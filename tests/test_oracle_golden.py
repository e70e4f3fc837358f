"""Pin oracle/oracle.py against golden vectors produced by the reference itself
(oracle/gen_golden.py).  Forward arithmetic must be BIT-EXACT: the oracle uses the same ATen CPU
primitives in the reference's order.  Gradients are compared with a tight tolerance because
ATen's reduction order may vary with the thread count."""
import numpy as np
import pytest
import torch

import cases
import oracle as O
from conftest import t


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else a.dtype)


def assert_bit_exact(a, b, what=""):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b) if a.dtype.kind == "f" else False)
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} elements differ"


def test_level_resolutions(golden):
    g = golden("resolutions")
    for key in [k for k in g if k.startswith("b")]:
        base, finest, L = (int(s[1:]) for s in key.split("_"))
        assert_bit_exact(O.level_resolutions(base, finest, L), g[key], key)
        assert_bit_exact(O.growth_factor(base, finest, L).reshape(()), g["growth_" + key], key)
    # the chair/lego default
    assert O.level_resolutions().tolist() == [16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203,
                                              256, 322, 406, 512]


def test_spatial_hash(golden):
    g = golden("hash")
    for log2T in (4, 10, 14, 19, 22, 24):
        assert_bit_exact(O.spatial_hash(t(g["coords"]), log2T), g[f"h{log2T}"])
        assert_bit_exact(O.spatial_hash_np(g["coords"], log2T), g[f"h{log2T}"])
    assert_bit_exact(O.spatial_hash(t(g["coords7"]), 19), g["h19_dim7"])
    assert_bit_exact(O.spatial_hash_np(g["coords7"], 19), g["h19_dim7"])


@pytest.mark.parametrize("name", ["hash_encode_unit_T10", "hash_encode_odd_T10", "hash_encode_odd_T19",
                                  "hash_encode_odd_T14_f1024"])
def test_hash_encode(golden, name):
    g = golden(name)
    log2T, L, finest = int(g["log2T"]), int(g["n_levels"]), int(g["finest"])
    tables = t(cases.synth_tables(L, log2T, 2)).requires_grad_(True)
    lo, hi = t(g["bbox"][0]), t(g["bbox"][1])
    res = O.level_resolutions(16, finest, L)
    out, keep, dbg = O.hash_encode(t(g["x"]), tables, lo, hi, res, log2T, return_debug=True)
    for l in range(L):
        assert_bit_exact(dbg[l]["hashed"].to(torch.int32), g["hashed"][l], f"hashed[{l}]")
        assert_bit_exact(dbg[l]["vmin"], g["vmin"][l], f"vmin[{l}]")
        assert_bit_exact(dbg[l]["vmax"], g["vmax"][l], f"vmax[{l}]")
    assert_bit_exact(out, g["out"], "features")
    assert_bit_exact(keep, g["keep"], "keep")
    assert keep.all()  # always True for L >= 2
    # backward: autograd of the restated forward vs the reference's autograd
    (out * t(g["dy"])).sum().backward()
    dense = np.zeros((L * (1 << log2T), 2), np.float32)
    dense[g["grad_rows"]] = g["grad_vals"]
    np.testing.assert_allclose(tables.grad.numpy().reshape(-1, 2), dense, rtol=1e-5, atol=1e-9)
    # and the analytic fp64 formula the CUDA backward is specified by
    an = O.hash_encode_grad_tables(t(g["x"]), t(g["dy"]), lo, hi, res, log2T, 2)
    np.testing.assert_allclose(an.numpy().reshape(-1, 2), dense, rtol=2e-5, atol=1e-7)


def test_hash_encode_single_level_mask(golden):
    g = golden("hash_encode_L1")
    tables = t(cases.synth_tables(1, 10, 2))
    res = torch.tensor([16.0])
    out, keep = O.hash_encode(t(g["x"]), tables, t(g["bbox"][0]), t(g["bbox"][1]), res, 10)
    assert_bit_exact(out, g["out"])
    assert_bit_exact(keep, g["keep"])
    assert not keep.all() and keep.any()


def test_sh(golden):
    g = golden("sh")
    for deg in (1, 2, 3, 4, 5):
        assert_bit_exact(O.sh_encode(t(g["dirs"]), deg), g[f"deg{deg}"], f"degree {deg}")


def test_mlp(golden):
    g = golden("mlp")
    ws = [t(g[f"w{i}"]).requires_grad_(True) for i in range(5)]
    x = t(g["x"]).requires_grad_(True)
    out = O.nerf_small(x, ws[:2], ws[2:])
    assert_bit_exact(out, g["out"])
    (out * t(g["dout"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g["dx"], rtol=1e-5, atol=1e-6)
    for i in range(5):
        np.testing.assert_allclose(ws[i].grad.numpy(), g[f"dw{i}"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag,white", [("black", False), ("white", True)])
def test_composite(golden, tag, white):
    g = golden("composite")
    raw = t(g["raw"]).requires_grad_(True)
    rgb, disp, acc, wts, depth, ent = O.composite(raw, t(g["z"]), t(g["rays_d"]), None, white)
    for name, val in (("rgb", rgb), ("disp", disp), ("acc", acc), ("weights", wts), ("depth", depth),
                      ("entropy", ent)):
        assert_bit_exact(val, g[f"{tag}_{name}"], name)
    assert np.isnan(g[f"{tag}_depth"][0])  # sum(w) == 0 -> NaN depth
    good = torch.ones(raw.shape[0], dtype=torch.bool)
    good[0] = False
    good[2] = False
    wm = t(g["w_misc"])
    loss = (rgb * t(g["w_rgb"]))[good].sum() + (acc * wm[:, 0])[good].sum() + (depth * wm[:, 1])[good].sum() \
        + (ent * wm[:, 2])[good].sum() + (wts * t(g["w_wts"]))[good].sum()
    loss.backward()
    np.testing.assert_allclose(raw.grad.numpy(), g[f"{tag}_draw"], rtol=1e-4, atol=1e-6)


def test_composite_noise(golden):
    g = golden("composite")
    rgb, _, _, wts, _, ent = O.composite(t(g["raw"]), t(g["z"]), t(g["rays_d"]), t(g["noise"]), True)
    assert_bit_exact(rgb, g["noise_rgb"])
    assert_bit_exact(wts, g["noise_weights"])
    assert_bit_exact(ent, g["noise_entropy"])


def test_sample_pdf(golden):
    g = golden("sample_pdf")
    assert_bit_exact(O.sample_pdf(t(g["bins"]), t(g["weights"]), t(g["u"])), g["samples_rand"])
    R, Ni = g["u"].shape
    assert_bit_exact(O.sample_pdf(t(g["bins"]), t(g["weights"]), O.det_u(R, Ni)), g["samples_det"])


def _render_case(g):
    log2T = int(g["log2T"])
    tables = (t(cases.synth_tables(16, log2T, 2)) * float(g["table_scale"])).requires_grad_(True)
    lo, hi = t(g["bbox"][0]), t(g["bbox"][1])
    res = O.level_resolutions()
    enc = lambda p: O.hash_encode(p, tables, lo, hi, res, log2T)
    cw = [t(g[f"coarse_w{i}"]).requires_grad_(True) for i in range(5)]
    fw = [t(g[f"fine_w{i}"]).requires_grad_(True) for i in range(5)]
    R, S, Ni = g["rays"].shape[0], int(g["N_samples"]), int(g["N_importance"])
    perturb, std = float(g["perturb"]), float(g["raw_noise_std"])
    # the reference's pytest=True draws (run_nerf_helpers.py:279-287, 531-534, 603-606)
    np.random.seed(0)
    t_rand = torch.Tensor(np.random.rand(R, S))
    np.random.seed(0)
    if perturb > 0:
        u = torch.Tensor(np.random.rand(R, Ni))
    else:  # pytest=True + det builds the linspace in fp64 numpy and casts (run_nerf_helpers.py:282-284)
        u = torch.Tensor(np.broadcast_to(np.linspace(0., 1., Ni), (R, Ni)).copy())
    n0 = n1 = None
    if std > 0:
        np.random.seed(0)
        n0 = torch.Tensor(np.random.rand(R, S) * std)
        np.random.seed(0)
        n1 = torch.Tensor(np.random.rand(R, S + Ni) * std)
    ret = O.render_rays(t(g["rays"]), enc, (cw[:2], cw[2:]), (fw[:2], fw[2:]), S, Ni, t_rand=t_rand, u=u,
                        noise0=n0, noise1=n1, white_bkgd=bool(g["white_bkgd"]), perturb=perturb)
    return ret, tables, cw, fw


@pytest.mark.parametrize("name", ["render_rays_perturb", "render_rays_det_noise"])
def test_render_rays(golden, name):
    g = golden(name)
    ret, tables, cw, fw = _render_case(g)
    for k in ("rgb_map", "depth_map", "acc_map", "sparsity_loss", "raw", "rgb0", "depth0", "acc0",
              "sparsity_loss0", "z_std"):
        assert_bit_exact(ret[k], g["ret_" + k], k)
    tgt = t(g["target"])
    loss = ((ret["rgb_map"] - tgt) ** 2).mean() + ((ret["rgb0"] - tgt) ** 2).mean() \
        + 1e-3 * (ret["sparsity_loss"].sum() + ret["sparsity_loss0"].sum())
    assert_bit_exact(loss.reshape(()), g["loss"])
    loss.backward()
    np.testing.assert_allclose(tables.grad.numpy(), g["grad_tables"], rtol=1e-4, atol=1e-9)
    for i in range(5):
        np.testing.assert_allclose(cw[i].grad.numpy(), g[f"coarse_dw{i}"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(fw[i].grad.numpy(), g[f"fine_dw{i}"], rtol=1e-4, atol=1e-7)


def test_tv_loss(golden):
    g = golden("tv_loss")
    tables = t(cases.synth_tables(16, 12, 2))
    for level in (0, 3, 7, 15):
        tab = tables[level].clone().requires_grad_(True)
        tv = O.total_variation(tab, 16, 512, level, 12, 16, t(g[f"l{level}_min_vertex"]))
        np.testing.assert_allclose(tv.detach().numpy(), g[f"l{level}_tv"], rtol=1e-6)
        tv.backward()
        dense = np.zeros((1 << 12, 2), np.float32)
        dense[g[f"l{level}_grad_rows"]] = g[f"l{level}_grad_vals"]
        np.testing.assert_allclose(tab.grad.numpy(), dense, rtol=1e-4, atol=1e-9)


def test_radam(golden):
    g = golden("radam")
    p, q = t(g["p0"]).clone(), t(g["q0"]).clone()
    mp, vp, mq, vq = (torch.zeros_like(a) for a in (p, p, q, q))
    lr = 0.01
    for step in range(g["grads_p"].shape[0]):
        O.radam_step(p, t(g["grads_p"][step]), mp, vp, step + 1, lr, 0.9, 0.99, 1e-8, 1e-6)
        O.radam_step(q, t(g["grads_q"][step]), mq, vq, step + 1, lr, 0.9, 0.99, 1e-15, 0.0)
        lr = 0.01 * (0.1 ** ((step + 1) / 10000.0))
        np.testing.assert_allclose(p.numpy(), g["traj_p"][step], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(q.numpy(), g["traj_q"][step], rtol=1e-6, atol=1e-12)


def test_rays(golden):
    g = golden("rays")
    H, W, focal = int(g["H"]), int(g["W"]), float(g["focal"])
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    o, d = O.pinhole_rays(H, W, K, t(g["c2w"]))
    assert_bit_exact(d, g["rays_d"])
    assert_bit_exact(o.contiguous(), g["rays_o"])
    of, df = O.pinhole_rays(H, W, K, t(g["c2w_f"]))
    no, nd = O.ndc_rays(H, W, K[0][0], 1., of.reshape(-1, 3), df.reshape(-1, 3))
    assert_bit_exact(no, g["ndc_o"])
    assert_bit_exact(nd, g["ndc_d"])
    # the package's CPU-side ray helpers (numpy variant and the torch branch used for CPU poses)
    from ray_util import get_rays, get_rays_np, get_ndc_rays
    o2, d2 = get_rays(H, W, K, t(g["c2w"]))
    assert_bit_exact(d2, g["rays_d"])
    on, dn = get_rays_np(H, W, K, g["c2w"])
    np.testing.assert_array_equal(dn, g["rays_d_np"])
    no2, nd2 = get_ndc_rays(H, W, K[0][0], 1., of.reshape(-1, 3), df.reshape(-1, 3))
    assert_bit_exact(no2, g["ndc_o"])
    assert_bit_exact(nd2, g["ndc_d"])


def test_reference_written_checkpoint_loads_into_our_modules(tmp_path):
    """A ``.tar`` written by the reference's OWN HashEmbedder / NeRFSmall / RAdam (run_nerf.py:663-680 format) loads
    into this repository's modules: same keys, shapes and values; the optimizer state keeps its moments and step
    counts.  (state_dict loading needs no GPU.)"""
    import sys
    import os
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_loader
    if not ref_loader.available():
        import pytest
        pytest.skip("reference tree not available here")
    ref = ref_loader.load("cpu")
    from embedding.hash_encoding import HashEmbedder
    from models import NeRFSmall
    from radam import RAdam
    box = (torch.tensor([-1.5, -1.5, -1.5]), torch.tensor([1.5, 1.5, 1.5]))
    torch.manual_seed(3)
    r_emb = ref.HashEmbedder(box, log2_hashmap_size=10)
    geo = dict(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, input_ch=32,
               input_ch_views=16)
    r_net, r_fine = ref.NeRFSmall(**geo), ref.NeRFSmall(**geo)
    groups = lambda emb, a, b: [{"params": list(a.parameters()) + list(b.parameters()), "weight_decay": 1e-6},
                                {"params": list(emb.parameters()), "eps": 1e-15}]
    r_opt = ref.RAdam(groups(r_emb, r_net, r_fine), lr=0.01, betas=(0.9, 0.99))
    for _ in range(3):   # give the optimizer real state
        y, _keep = r_emb(torch.rand(50, 3) * 3 - 1.5)
        out = r_net(torch.cat([y, torch.rand(50, 16)], -1)) + r_fine(torch.cat([y, torch.rand(50, 16)], -1))
        r_opt.zero_grad()
        out.square().mean().backward()
        r_opt.step()
    path = os.path.join(tmp_path, "000003.tar")
    torch.save({"global_step": 3, "network_fn_state_dict": r_net.state_dict(),
                "network_fine_state_dict": r_fine.state_dict(), "embed_fn_state_dict": r_emb.state_dict(),
                "optimizer_state_dict": r_opt.state_dict()}, path)

    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    emb = HashEmbedder(box, log2_hashmap_size=10)
    net, fine = NeRFSmall(**geo), NeRFSmall(**geo)
    opt = RAdam(groups(emb, net, fine), lr=0.01, betas=(0.9, 0.99))
    opt._span_cache = ("stale", [])          # a plan from before the resume must not survive it
    emb.load_state_dict(ckpt["embed_fn_state_dict"])
    net.load_state_dict(ckpt["network_fn_state_dict"])
    fine.load_state_dict(ckpt["network_fine_state_dict"])
    opt.load_state_dict(ckpt["optimizer_state_dict"])
    assert "_span_cache" not in opt.__dict__
    for a, b in zip(list(emb.parameters()) + list(net.parameters()) + list(fine.parameters()),
                    list(r_emb.parameters()) + list(r_net.parameters()) + list(r_fine.parameters())):
        assert a.shape == b.shape and torch.equal(a.detach(), b.detach())
    for ours_p, ref_p in zip([p for g in opt.param_groups for p in g["params"]],
                             [p for g in r_opt.param_groups for p in g["params"]]):
        so, sr = opt.state[ours_p], r_opt.state[ref_p]
        assert so["step"] == sr["step"] == 3
        assert torch.equal(so["exp_avg"], sr["exp_avg"]) and torch.equal(so["exp_avg_sq"], sr["exp_avg_sq"])
    assert opt.param_groups[0]["weight_decay"] == 1e-6 and opt.param_groups[1]["eps"] == 1e-15
    # and the other way round: what we save has the reference's keys and loads into the reference's modules
    r_emb.load_state_dict(emb.state_dict())
    r_net.load_state_dict(net.state_dict())
    r_opt.load_state_dict(opt.state_dict())
