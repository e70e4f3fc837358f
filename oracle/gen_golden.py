"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions on seeded inputs.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

The reference has no golden vectors of its own, so these files -- outputs of the
reference itself, executed through oracle/ref_loader.py -- are what pins oracle/oracle.py and,
through it, the CUDA path.  Each .npz stores the inputs next to the outputs so the tests need
neither /root/reference nor this script at run time.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
warnings.filterwarnings("ignore")
torch.set_num_threads(1)  # deterministic accumulation order inside ATen reductions


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"  wrote {os.path.relpath(path)}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def make_embedder(ref, bbox, log2T, finest=512, n_levels=16, F=2, base=16):
    box = (torch.tensor(bbox[0], dtype=torch.float32), torch.tensor(bbox[1], dtype=torch.float32))
    emb = ref.HashEmbedder(box, n_levels=n_levels, n_features_per_level=F, log2_hashmap_size=log2T,
                           base_resolution=base, finest_resolution=finest)
    tables = cases.synth_tables(n_levels, log2T, F)
    for l in range(n_levels):
        emb.embeddings[l].weight.data.copy_(T(tables[l]))
    return emb, tables


def make_mlp(ref, seed):
    sig, col = cases.mlp_weights(seed)
    net = ref.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3,
                        hidden_dim_color=64, input_ch=32, input_ch_views=16)
    for lin, w in zip(net.sigma_net, sig):
        lin.weight.data.copy_(T(w))
    for lin, w in zip(net.color_net, col):
        lin.weight.data.copy_(T(w))
    return net, sig, col


# ----------------------------------------------------------------------------------------------
def gen_resolutions(ref):
    rows = {}
    for base, finest, L in [(16, 512, 16), (16, 1024, 16), (16, 2048, 16), (16, 512, 8),
                            (4, 64, 6), (16, 4096, 16), (8, 300, 12), (16, 512, 2)]:
        box = (torch.zeros(3), torch.ones(3))
        emb = ref.HashEmbedder(box, n_levels=L, log2_hashmap_size=4, base_resolution=base,
                               finest_resolution=finest)
        res = [float(torch.floor(emb.base_resolution * emb.b ** i)) for i in range(L)]
        rows[f"b{base}_f{finest}_L{L}"] = np.array(res, dtype=np.float32)
        rows[f"growth_b{base}_f{finest}_L{L}"] = np.array(float(emb.b), dtype=np.float32)
    save("resolutions", **rows)


def gen_hash(ref):
    rs = np.random.RandomState(7)
    coords = rs.randint(0, 2050, size=(4096, 3)).astype(np.int64)
    coords[:8] = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [512, 512, 512],
                           [513, 513, 513], [2049, 2049, 2049], [1, 1, 1]])
    out = {"coords": coords}
    for log2T in (4, 10, 14, 19, 22, 24):
        out[f"h{log2T}"] = ref.hash(T(coords)[None], log2T)[0].numpy()
    c7 = rs.randint(0, 300, size=(64, 7)).astype(np.int64)
    out["coords7"] = c7
    out["h19_dim7"] = ref.hash(T(c7)[None], 19)[0].numpy()
    save("hash", **out)


def gen_hash_encode(ref, name, bbox, log2T, n, seed, finest=512, L=16):
    emb, tables = make_embedder(ref, bbox, log2T, finest=finest, n_levels=L)
    x = cases.points_in_box(n, bbox, seed)
    xt = T(x)
    # per-level internals, walking the levels exactly as forward() does (clamp persists)
    emb.xyz = xt
    idx_all, vmin_all, vmax_all = [], [], []
    for i in range(L):
        res = torch.floor(emb.base_resolution * emb.b ** i)
        vmin, vmax, hashed, keep3 = emb.get_voxel_vertices(res)
        idx_all.append(hashed.numpy().astype(np.int32))
        vmin_all.append(vmin.numpy())
        vmax_all.append(vmax.numpy())
    # the real forward + backward
    out, keep = emb(xt)
    dy = np.random.RandomState(seed + 1).randn(n, L * 2).astype(np.float32)
    (out * T(dy)).sum().backward()
    grads = np.stack([emb.embeddings[l].weight.grad.numpy() for l in range(L)])
    nz = np.nonzero(np.abs(grads).sum(-1).reshape(-1))[0]
    save(name, bbox=np.array(bbox, dtype=np.float32), log2T=log2T, finest=finest, n_levels=L,
         x=x, out=out.detach().numpy(), keep=keep.numpy(), hashed=np.stack(idx_all),
         vmin=np.stack(vmin_all), vmax=np.stack(vmax_all), dy=dy,
         grad_rows=nz.astype(np.int64), grad_vals=grads.reshape(-1, 2)[nz])


def gen_keep_mask_single_level(ref):
    """n_levels == 1 is the only case where the returned keep mask can be False."""
    bbox = cases.BBOX_ODD
    box = (torch.tensor(bbox[0]), torch.tensor(bbox[1]))
    x = cases.points_in_box(64, bbox, 3)
    emb = ref.HashEmbedder(box, n_levels=1, log2_hashmap_size=10, base_resolution=16, finest_resolution=512)
    # n_levels=1 makes b = exp(x/0) = inf/nan; level 0 uses b**0 = 1 regardless.
    tables = cases.synth_tables(1, 10, 2)
    emb.embeddings[0].weight.data.copy_(T(tables[0]))
    out, keep = emb(T(x))
    save("hash_encode_L1", bbox=np.array(bbox, dtype=np.float32), x=x, out=out.detach().numpy(),
         keep=keep.numpy())


def gen_sh(ref):
    rs = np.random.RandomState(11)
    d = rs.randn(512, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d[0] = (0, 0, 1)
    d[1] = (1, 0, 0)
    d[2] = (0, -1, 0)
    d[3] = (0.57735026, 0.57735026, 0.57735026)
    d[4:8] *= np.float32(1.7)  # unnormalised input is legal: SH is just polynomials
    out = {"dirs": d}
    for deg in (1, 2, 3, 4, 5):
        out[f"deg{deg}"] = ref.SHEncoder(degree=deg)(T(d)).numpy()
    save("sh", **out)


def gen_mlp(ref):
    net, sig, col = make_mlp(ref, 21)
    rs = np.random.RandomState(22)
    x = rs.randn(384, 48).astype(np.float32)
    x[:, :32] *= 0.3
    xt = T(x).requires_grad_(True)
    out = net(xt)
    dout = rs.randn(384, 4).astype(np.float32)
    (out * T(dout)).sum().backward()
    save("mlp", x=x, dout=dout, out=out.detach().numpy(), dx=xt.grad.numpy(),
         w0=sig[0], w1=sig[1], w2=col[0], w3=col[1], w4=col[2],
         dw0=net.sigma_net[0].weight.grad.numpy(), dw1=net.sigma_net[1].weight.grad.numpy(),
         dw2=net.color_net[0].weight.grad.numpy(), dw3=net.color_net[1].weight.grad.numpy(),
         dw4=net.color_net[2].weight.grad.numpy())


def gen_composite(ref):
    rs = np.random.RandomState(31)
    R, S = 96, 48
    raw = rs.randn(R, S, 4).astype(np.float32)
    raw[..., 3] = raw[..., 3] * 3.0 + 0.5
    raw[0, :, 3] = -1.0                      # sigma <= 0 everywhere: sum w = 0 -> NaN depth (A.1.6)
    raw[1, :, 3] = 1e4                       # saturated alpha
    raw[2, :, 3] = 0.0
    z = np.sort(2.0 + 4.0 * rs.rand(R, S).astype(np.float32), axis=-1)
    z[3, 10:14] = z[3, 10]                   # duplicate depths after resampling
    d = rs.randn(R, 3).astype(np.float32)
    w_rgb = rs.randn(R, 3).astype(np.float32)
    w_misc = rs.randn(R, 4).astype(np.float32)
    w_wts = rs.randn(R, S).astype(np.float32)
    out = dict(raw=raw, z=z, rays_d=d, w_rgb=w_rgb, w_misc=w_misc, w_wts=w_wts)
    for tag, white in (("black", False), ("white", True)):
        rt = T(raw).requires_grad_(True)
        rgb, disp, acc, wts, depth, ent = ref.raw2outputs(rt, T(z), T(d), 0, white)
        # rows 0 and 2 have sum(w)=0 (NaN depth/disp); keep them out of the scalar loss
        good = torch.ones(R, dtype=torch.bool)
        good[0] = False
        good[2] = False
        loss = (rgb * T(w_rgb))[good].sum() + (acc * T(w_misc[:, 0]))[good].sum() \
            + (depth * T(w_misc[:, 1]))[good].sum() + (ent * T(w_misc[:, 2]))[good].sum() \
            + (wts * T(w_wts))[good].sum()
        loss.backward()
        out.update({f"{tag}_rgb": rgb.detach().numpy(), f"{tag}_disp": disp.detach().numpy(),
                    f"{tag}_acc": acc.detach().numpy(), f"{tag}_weights": wts.detach().numpy(),
                    f"{tag}_depth": depth.detach().numpy(), f"{tag}_entropy": ent.detach().numpy(),
                    f"{tag}_draw": rt.grad.numpy()})
    # noise path with pytest=True: noise = np.random.rand(R,S) * std after seed(0)  (:603-606)
    rgb, disp, acc, wts, depth, ent = ref.raw2outputs(T(raw), T(z), T(d), 0.5, True, pytest=True)
    np.random.seed(0)
    noise = (np.random.rand(R, S) * 0.5)
    out.update(noise=torch.Tensor(noise).numpy(), noise_rgb=rgb.numpy(), noise_weights=wts.numpy(),
               noise_entropy=ent.numpy())
    save("composite", **out)


def gen_sample_pdf(ref):
    rs = np.random.RandomState(41)
    R, S, Ni = 64, 48, 96
    z = np.sort(2.0 + 4.0 * rs.rand(R, S).astype(np.float32), axis=-1)
    mids = (0.5 * (z[:, 1:] + z[:, :-1])).astype(np.float32)
    w = (rs.rand(R, S - 2).astype(np.float32)) ** 4
    w[0] = 0.0                               # flat pdf
    w[1] = 0.0
    w[1, 17] = 1.0                           # a delta
    w[2, :20] = 0.0                          # long zero-ish runs: denom < 1e-5 branch
    u = rs.rand(R, Ni).astype(np.float32)
    u[3, :4] = (0.0, 1.0 - 2 ** -24, 0.5, 2 ** -30)
    # the reference draws u internally; patch torch.rand for the call to inject ours
    real_rand = torch.rand
    torch.rand = lambda *a, **k: T(u).clone()
    try:
        rnd = ref.sample_pdf(T(mids), T(w), Ni, det=False)
    finally:
        torch.rand = real_rand
    det = ref.sample_pdf(T(mids), T(w), Ni, det=True)
    save("sample_pdf", bins=mids, weights=w, u=u, samples_rand=rnd.numpy(), samples_det=det.numpy())


def gen_render_rays(ref, name, bbox, perturb, white, noise_std, R=48, S=24, Ni=40, log2T=12):
    emb, tables = make_embedder(ref, bbox, log2T)
    # larger table values so that densities are not all ~0 and the scene is non-trivial
    for l in range(16):
        emb.embeddings[l].weight.data.mul_(3000.0)
    coarse, sig0, col0 = make_mlp(ref, 51)
    fine, sig1, col1 = make_mlp(ref, 52)
    sh = ref.SHEncoder()
    rays = cases.rays(R, 53)
    qfn = lambda inputs, viewdirs, fn: ref.run_network(inputs, viewdirs, fn, embed_fn=emb,
                                                       embeddirs_fn=sh, netchunk=1 << 16)
    # record what the reference's own sample_pdf saw and returned (run_nerf_helpers.py:548): the GPU tests feed
    # these z_samples into the fine pass so that it can be held to the strict tolerance (the resampling itself is
    # checked against a per-sample conditioning bound, tests/test_gpu_parity.py::sample_pdf_tolerance)
    real_sample_pdf = ref.helpers.sample_pdf
    spied = {}

    def spy(bins, weights, N_samples, det=False, pytest=False):
        out = real_sample_pdf(bins, weights, N_samples, det=det, pytest=pytest)
        spied.update(bins=bins.detach().clone(), weights=weights.detach().clone(), samples=out.detach().clone())
        return out

    ref.helpers.sample_pdf = spy
    try:
        ret = ref.render_rays(T(rays), coarse, qfn, S, embed_fn=emb, retraw=True, perturb=perturb,
                              N_importance=Ni, network_fine=fine, white_bkgd=white,
                              raw_noise_std=noise_std, pytest=True)
    finally:
        ref.helpers.sample_pdf = real_sample_pdf
    rs = np.random.RandomState(54)
    target = rs.rand(R, 3).astype(np.float32)
    loss = ((ret["rgb_map"] - T(target)) ** 2).mean() + ((ret["rgb0"] - T(target)) ** 2).mean() \
        + 1e-3 * (ret["sparsity_loss"].sum() + ret["sparsity_loss0"].sum())
    loss.backward()
    gt = np.stack([emb.embeddings[l].weight.grad.numpy() for l in range(16)])
    out = dict(bbox=np.array(bbox, dtype=np.float32), log2T=log2T, table_scale=3000.0, rays=rays,
               N_samples=S, N_importance=Ni, perturb=perturb, white_bkgd=white, raw_noise_std=noise_std,
               target=target, loss=loss.detach().numpy(), grad_tables=gt)
    for k, v in ret.items():
        out["ret_" + k] = v.detach().numpy()
    out["pdf_bins"] = spied["bins"].numpy()
    out["pdf_weights"] = spied["weights"].numpy()
    out["z_samples"] = spied["samples"].numpy()
    for tag, net in (("coarse", coarse), ("fine", fine)):
        for i, lin in enumerate(list(net.sigma_net) + list(net.color_net)):
            out[f"{tag}_w{i}"] = lin.weight.detach().numpy()
            out[f"{tag}_dw{i}"] = lin.weight.grad.numpy()
    save(name, **out)


def gen_tv(ref):
    emb, tables = make_embedder(ref, cases.BBOX_UNIT, 12)
    out = {"log2T": 12}
    for level in (0, 3, 7, 15):
        torch.manual_seed(100 + level)
        # capture the random origin the reference draws (loss.py:25) by replaying the generator
        real_randint = torch.randint
        grabbed = {}

        def spy(*a, **k):
            v = real_randint(*a, **k)
            grabbed["v"] = v.clone()
            return v

        torch.randint = spy
        try:
            w = emb.embeddings[level].weight
            w.grad = None
            tv = ref.total_variation_loss(emb.embeddings[level], 16, 512, level, 12, n_levels=16)
            tv.backward()
        finally:
            torch.randint = real_randint
        nz = np.nonzero(np.abs(w.grad.numpy()).sum(-1))[0]
        out[f"l{level}_min_vertex"] = grabbed["v"].numpy()
        out[f"l{level}_tv"] = tv.detach().numpy()
        out[f"l{level}_grad_rows"] = nz.astype(np.int64)
        out[f"l{level}_grad_vals"] = w.grad.numpy()[nz]
    save("tv_loss", **out)


def gen_radam(ref):
    rs = np.random.RandomState(61)
    p0 = rs.randn(257, 3).astype(np.float32) * 0.1
    q0 = rs.randn(64, 2).astype(np.float32) * 1e-4
    p = torch.nn.Parameter(T(p0).clone())
    q = torch.nn.Parameter(T(q0).clone())
    opt = ref.RAdam([{"params": [p], "weight_decay": 1e-6}, {"params": [q], "eps": 1e-15}],
                    lr=0.01, betas=(0.9, 0.99))
    grads_p, grads_q, traj_p, traj_q = [], [], [], []
    for step in range(12):
        gp = rs.randn(257, 3).astype(np.float32)
        gq = (rs.randn(64, 2) * 1e-3).astype(np.float32)
        p.grad = T(gp).clone()
        q.grad = T(gq).clone()
        opt.step()
        lr = 0.01 * (0.1 ** ((step + 1) / 10000.0))
        for g in opt.param_groups:
            g["lr"] = lr
        grads_p.append(gp)
        grads_q.append(gq)
        traj_p.append(p.detach().numpy().copy())
        traj_q.append(q.detach().numpy().copy())
    save("radam", p0=p0, q0=q0, grads_p=np.stack(grads_p), grads_q=np.stack(grads_q),
         traj_p=np.stack(traj_p), traj_q=np.stack(traj_q))


def gen_rays(ref):
    rs = np.random.RandomState(71)
    H, W, focal = 23, 31, 27.5
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    q, _ = np.linalg.qr(rs.randn(3, 3))
    c2w = np.concatenate([q, rs.randn(3, 1)], -1).astype(np.float32)
    o, d = ref.get_rays(H, W, K, T(c2w))
    o_np, d_np = ref.get_rays_np(H, W, K, c2w)
    # forward-facing pose for the NDC warp (looking down -z from z > 0)
    c2w_f = np.array([[1, 0, 0, 0.05], [0, 1, 0, -0.02], [0, 0, 1, 0.3]], np.float32)
    of, df = ref.get_rays(H, W, K, T(c2w_f))
    no, nd = ref.get_ndc_rays(H, W, K[0][0], 1., of.reshape(-1, 3), df.reshape(-1, 3))
    save("rays", H=H, W=W, focal=focal, c2w=c2w, rays_o=o.contiguous().numpy(), rays_d=d.numpy(),
         rays_o_np=np.ascontiguousarray(o_np), rays_d_np=d_np, c2w_f=c2w_f, ndc_o=no.numpy(), ndc_d=nd.numpy())


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not found; golden vectors can only be regenerated where "
                         "/root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load("cpu")
    print("generating golden vectors from the reference at", ref_loader.REF_ROOT)
    gen_resolutions(ref)
    gen_hash(ref)
    gen_hash_encode(ref, "hash_encode_unit_T10", cases.BBOX_UNIT, 10, 384, 101)
    gen_hash_encode(ref, "hash_encode_odd_T10", cases.BBOX_ODD, 10, 384, 102)
    gen_hash_encode(ref, "hash_encode_odd_T19", cases.BBOX_ODD, 19, 256, 103)
    gen_hash_encode(ref, "hash_encode_odd_T14_f1024", cases.BBOX_ODD, 14, 256, 104, finest=1024)
    gen_keep_mask_single_level(ref)
    gen_sh(ref)
    gen_mlp(ref)
    gen_composite(ref)
    gen_sample_pdf(ref)
    gen_render_rays(ref, "render_rays_perturb", cases.BBOX_UNIT, 1.0, True, 0.0)
    gen_render_rays(ref, "render_rays_det_noise", cases.BBOX_ODD, 0.0, False, 1.0)
    gen_tv(ref)
    gen_radam(ref)
    gen_rays(ref)


if __name__ == "__main__":
    main()
