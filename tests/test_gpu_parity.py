"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in modules) against
 (a) golden vectors produced by the reference itself and (b) the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): hash indices / voxel vertices / encoded features bit-exact;
floating-point outputs rel 1e-5 forward; atomically accumulated gradients rel 1e-4."""
import numpy as np
import pytest
import torch

import cases
import oracle as O
from conftest import t

pytestmark = pytest.mark.gpu
DEV = "cuda"

FWD_RTOL, GRAD_RTOL = 1e-5, 1e-4


def g32(a):
    return t(np.asarray(a, dtype=np.float32)).to(DEV)


def bit_equal(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype == np.float32:
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    assert same.all(), f"{np.count_nonzero(~same)} of {same.size} elements differ"


def close(a, b, rtol, atol=0.0, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True, err_msg=what)


def samples_within_bound(got, bins, weights, u, what="", cdf_eps=1e-6):
    """EVERY resampled depth must lie within the analytic bound of oracle.sample_pdf_tolerance: rtol 1e-5 /
    atol 2e-5 plus the forward error of a CDF that carries 1e-6 of summation-order noise (bin width x 2e-6 / denom),
    and anywhere inside the neighbouring bins only for the branch-chaotic samples (denom within 2e-6 of the
    reference's `denom < 1e-5` switch, run_nerf_helpers.py:302, or u within 1e-6 of a CDF entry).  Pipeline-level
    callers, whose pdf weights come from two coarse passes that agree to 5e-5 rather than bit for bit, pass
    cdf_eps = 1e-5.  Returns the fraction of branch-chaotic samples."""
    got = got.detach().cpu() if isinstance(got, torch.Tensor) else torch.from_numpy(np.asarray(got))
    as_t = lambda a: a if isinstance(a, torch.Tensor) else t(a)
    want, tol, chaotic = O.sample_pdf_tolerance(as_t(bins), as_t(weights), as_t(u), rtol=FWD_RTOL, atol=2e-5,
                                                cdf_eps=cdf_eps)
    err = (got - want).abs()
    bad = ~(err <= tol)
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())} of {bad.numel()} samples outside their bound; worst "
                                 f"err {float(err[bad].max()):.3e} vs tol {float(tol[bad][err[bad].argmax()]):.3e}")
    return float(chaotic.float().mean())


def recorded_resample_within_bound(call, what=""):
    """A recorded ops.resample call (z_vals, weights, n, u, u_det, samples): the oracle's sample_pdf on the SAME
    inputs (run_nerf_helpers.py:547-549) bounds every sample we drew."""
    z, w, n, u, u_det, samples = call
    z, w = z.detach().cpu(), w.detach().cpu()
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    uu = u.detach().cpu() if u is not None else u_det.detach().cpu().expand(z.shape[0], n).contiguous()
    return samples_within_bound(samples, mids, w[:, 1:-1].contiguous(), uu, what=what)


def make_embedder(bbox, log2T, finest=512, L=16, F=2, scale=1.0):
    from embedding.hash_encoding import HashEmbedder
    box = (torch.tensor(bbox[0], dtype=torch.float32), torch.tensor(bbox[1], dtype=torch.float32))
    emb = HashEmbedder(box, n_levels=L, n_features_per_level=F, log2_hashmap_size=log2T, finest_resolution=finest)
    tables = cases.synth_tables(L, log2T, F) * np.float32(scale)
    with torch.no_grad():
        for l in range(L):
            emb.embeddings[l].weight.copy_(t(tables[l]))
    return emb.to(DEV), tables


def make_mlp(weights5):
    from models import NeRFSmall
    net = NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                    input_ch=32, input_ch_views=16)
    with torch.no_grad():
        for lin, w in zip(list(net.sigma_net) + list(net.color_net), weights5):
            lin.weight.copy_(t(np.asarray(w)))
    return net.to(DEV)


# ---------------------------------------------------------------------------------------------- hash
def test_spatial_hash(golden):
    from embedding.hash_encoding import hash
    g = golden("hash")
    for log2T in (4, 10, 14, 19, 22, 24):
        bit_equal(hash(t(g["coords"]).to(DEV), log2T), g[f"h{log2T}"])
    bit_equal(hash(t(g["coords7"]).to(DEV), 19), g["h19_dim7"])
    # leading batch dims, like hash(voxel_indices[N,8,3]) in the reference
    c = t(g["coords"]).to(DEV).reshape(64, 64, 3)
    bit_equal(hash(c, 19), g["h19"].reshape(64, 64))


@pytest.mark.parametrize("name", ["hash_encode_unit_T10", "hash_encode_odd_T10", "hash_encode_odd_T19",
                                  "hash_encode_odd_T14_f1024"])
def test_hash_encode_golden(golden, name):
    g = golden(name)
    log2T, L, finest = int(g["log2T"]), int(g["n_levels"]), int(g["finest"])
    emb, _ = make_embedder(g["bbox"], log2T, finest=finest, L=L)
    x = g32(g["x"])
    hashed, vmin, vmax = emb.voxel_vertices(x)
    bit_equal(hashed.to(torch.int32), g["hashed"])            # indices: bit-exact
    bit_equal(vmin, g["vmin"])                                # voxel vertices: bit-exact
    bit_equal(vmax, g["vmax"])
    out, keep = emb(x)
    bit_equal(out, g["out"])                                  # encoded features: bit-exact
    bit_equal(keep, g["keep"])
    (out * g32(g["dy"])).sum().backward()
    dense = np.zeros((L * (1 << log2T), 2), np.float32)
    dense[g["grad_rows"]] = g["grad_vals"]
    got = torch.stack([e.weight.grad for e in emb.embeddings]).reshape(-1, 2)
    close(got, dense, GRAD_RTOL, atol=1e-7, what="table gradients")


def test_hash_encode_single_level_keep_mask(golden):
    g = golden("hash_encode_L1")
    emb, _ = make_embedder(g["bbox"], 10, L=1)
    out, keep = emb(g32(g["x"]))
    bit_equal(out, g["out"])
    bit_equal(keep, g["keep"])
    assert not bool(keep.all()) and bool(keep.any())


@pytest.mark.parametrize("log2T,bbox,n", [(14, cases.BBOX_ODD, 40_000), (19, cases.BBOX_UNIT, 40_000),
                                           (19, cases.BBOX_ODD, 33_333), (22, cases.BBOX_ODD, 8_191)])
def test_hash_encode_vs_oracle(log2T, bbox, n):
    emb, tables = make_embedder(bbox, log2T)
    x = cases.points_in_box(n, bbox, seed=1000 + log2T)
    lo, hi = t(np.float32(bbox[0])), t(np.float32(bbox[1]))
    want, want_keep = O.hash_encode(t(x), t(tables), lo, hi, O.level_resolutions(), log2T)
    out, keep = emb(g32(x))
    bit_equal(out, want)
    bit_equal(keep, want_keep)
    dy = np.random.RandomState(5).randn(n, 32).astype(np.float32)
    (out * g32(dy)).sum().backward()
    an = O.hash_encode_grad_tables(t(x), t(dy), lo, hi, O.level_resolutions(), log2T, 2)
    got = torch.stack([e.weight.grad for e in emb.embeddings]).double().cpu()
    scale = an.abs().max().item()
    assert (got - an).abs().max().item() <= GRAD_RTOL * scale


@pytest.mark.parametrize("lpg", [1, 2, 4, 8, 16])
def test_hash_encode_launch_shapes_agree(lpg):
    from hn_b200 import _lib
    emb, tables = make_embedder(cases.BBOX_ODD, 12)
    x = g32(cases.points_in_box(5000, cases.BBOX_ODD, 77))
    ref_out, _ = emb(x)
    _lib.set_tuning("hash_fwd_lpg", lpg)
    _lib.set_tuning("hash_bwd_lpg", lpg)
    try:
        out, _ = emb(x)
        bit_equal(out, ref_out)
        dy = torch.randn_like(out)
        (out * dy).sum().backward()
        g1 = torch.stack([e.weight.grad for e in emb.embeddings]).clone()
    finally:
        _lib.set_tuning("hash_fwd_lpg", 0)
        _lib.set_tuning("hash_bwd_lpg", 0)
    for e in emb.embeddings:
        e.weight.grad = None
    out2, _ = emb(x)
    (out2 * dy).sum().backward()
    g2 = torch.stack([e.weight.grad for e in emb.embeddings])
    close(g1, g2, GRAD_RTOL, atol=1e-6 * g2.abs().max().item())


@pytest.mark.parametrize("log2T,bbox,n,agg", [(14, cases.BBOX_ODD, 50_000, -1), (19, cases.BBOX_UNIT, 33_333, -1),
                                               (12, cases.BBOX_ODD, 257, -1), (19, cases.BBOX_ODD, 40_000, 1)])
def test_hash_encode_coherent_path_vs_oracle(log2T, bbox, n, agg):
    """Counting sort + sorted gather + warp-aggregated scatter: forward still bit-exact, gradients within the
    atomic tolerance; also the aggregated scatter on caller-ordered points (agg=1)."""
    from hn_b200 import _lib, ops
    emb, tables = make_embedder(bbox, log2T)
    emb.coherent = True if agg < 0 else False
    x = cases.points_in_box(n, bbox, seed=2000 + log2T)
    if agg > 0:  # ray-like ordering: consecutive points are close, so runs exist without sorting
        x = x[np.lexsort((x[:, 0], x[:, 1], x[:, 2]))]
    lo, hi = t(np.float32(bbox[0])), t(np.float32(bbox[1]))
    want, want_keep = O.hash_encode(t(x), t(tables), lo, hi, O.level_resolutions(), log2T)
    _lib.set_tuning("hash_bwd_agg", agg)
    try:
        out, keep = emb(g32(x))
        bit_equal(out, want)
        bit_equal(keep, want_keep)
        dy = np.random.RandomState(5).randn(n, 32).astype(np.float32)
        (out * g32(dy)).sum().backward()
    finally:
        _lib.set_tuning("hash_bwd_agg", -1)
    an = O.hash_encode_grad_tables(t(x), t(dy), lo, hi, O.level_resolutions(), log2T, 2)
    got = torch.stack([e.weight.grad for e in emb.embeddings]).double().cpu()
    assert (got - an).abs().max().item() <= GRAD_RTOL * an.abs().max().item()
    # the sort is a permutation of the rows and keeps the coordinates intact
    xs4 = ops.hash_sort_points(g32(x), emb._geometry(torch.device(DEV))[0], 32)
    rows = xs4[:, 3].contiguous().view(torch.int32).long()
    assert torch.equal(torch.sort(rows).values, torch.arange(n, device=DEV))
    bit_equal(xs4[:, :3], g32(x)[rows])


@pytest.mark.parametrize("bbox", [cases.BBOX_ODD, cases.BBOX_UNIT,
                                  ((-3e-21, -1e-21, -2e-21), (1e-21, 4e-21, 2e-21)),      # cell sizes below 2^-60
                                  ((-7.3e5, -1.1e6, -2.0e5), (9.1e5, 4.4e5, 8.8e5)),
                                  ((0.0, 0.0, 0.0), (1.0, 3.0, 7.0))])
@pytest.mark.parametrize("coherent", [False, True])
def test_hash_encode_cell_boundaries_bit_exact(bbox, coherent):
    """Points sitting exactly on (and one ulp either side of) voxel faces of every level: the quotient
    (x - min) / g is within an ulp of an integer there, so any division that is not correctly rounded moves a
    point into the neighbouring voxel.  Pins the hoisted-reciprocal division against the reference's true
    division, on both sides of its range switch (LevelGeom::fast)."""
    log2T = 14
    emb, tables = make_embedder(bbox, log2T)
    emb.coherent = coherent
    lo32, hi32 = np.float32(bbox[0]), np.float32(bbox[1])
    res = O.level_resolutions().numpy().astype(np.float32)
    rs = np.random.RandomState(99)
    pts = []
    for l in range(16):
        g = (hi32 - lo32) / res[l]                                   # fp32, as the reference computes it
        k = rs.randint(0, int(res[l]) + 1, size=(600, 3)).astype(np.float32)
        base = (k * g + lo32).astype(np.float32)
        pts += [base, np.nextafter(base, np.float32(np.inf)), np.nextafter(base, np.float32(-np.inf))]
    tiny = np.float32(bbox[0]) + np.float32([1e-45, 1e-40, 1e-38]) * rs.rand(64, 3).astype(np.float32)
    x = np.concatenate(pts + [tiny.astype(np.float32)]).astype(np.float32)
    lo, hi = t(lo32), t(hi32)
    want, want_keep = O.hash_encode(t(x), t(tables), lo, hi, O.level_resolutions(), log2T)
    out, keep = emb(g32(x))
    bit_equal(out, want)
    bit_equal(keep, want_keep)
    # the stand-alone voxel kernel (true division per point) and the encoder's hoisted division agree on cells
    hashed, vmin, vmax = emb.voxel_vertices(g32(x))
    xyz = t(x)
    for l in range(16):
        xyz, _, wmin, wmax, whash, _ = O.voxel_vertices(xyz, lo, hi, O.level_resolutions()[l], log2T)
        bit_equal(vmin[l], wmin)
        bit_equal(hashed[l], whash)


@pytest.mark.parametrize("parts", [[(0, 4), (4, 8), (8, 12), (12, 16)], [(0, 16)], [(0, 6), (6, 7), (7, 16)],
                                   [(0, 8), (8, 8), (8, 16)]])
def test_hash_encode_backward_level_buckets(parts):
    """hn_hash_encode_bwd_sorted_levels over a partition of the levels == one full scatter (the gradient buckets
    of the data-parallel path), including ranges that are not multiples of the 4-levels-per-thread mapping."""
    from hn_b200 import ops
    log2T, n = 14, 30_000
    emb, tables = make_embedder(cases.BBOX_ODD, log2T)
    x = g32(cases.points_in_box(n, cases.BBOX_ODD, seed=4242))
    dy = g32(np.random.RandomState(6).randn(n, 32).astype(np.float32))
    box, res = emb._geometry(torch.device(DEV))
    xs4 = ops.hash_sort_points(x, box, 32)
    full = torch.zeros(16 << log2T, 2, device=DEV)
    ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, full)
    got = torch.zeros_like(full)
    for b, e in parts:
        before = got.clone()
        ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, got, levels=(b, e))
        changed = (got != before).view(16, -1).any(dim=1).cpu().numpy()
        assert not changed[:b].any() and not changed[e:].any(), "a bucket wrote outside its levels"
    scale = full.abs().max().item()
    assert (got - full).abs().max().item() <= GRAD_RTOL * scale
    with pytest.raises(RuntimeError):
        ops.hash_encode_backward_sorted(xs4, dy, box, res, 16, 2, log2T, got, levels=(5, 17))


@pytest.mark.parametrize("n", [0, 1, 255, 257])
def test_hash_encode_ragged_and_empty(n):
    emb, tables = make_embedder(cases.BBOX_ODD, 10)
    x = cases.points_in_box(max(n, 16), cases.BBOX_ODD, 9)[:n]
    for coherent in (False, True):
        emb.coherent = coherent
        out, keep = emb(g32(x).reshape(n, 3))
        assert out.shape == (n, 32) and keep.shape == (n,)
    if n:
        want, _ = O.hash_encode(t(x), t(tables), t(np.float32(cases.BBOX_ODD[0])), t(np.float32(cases.BBOX_ODD[1])),
                                O.level_resolutions(), 10)
        bit_equal(out, want)


@pytest.mark.parametrize("F,L", [(4, 8), (1, 5), (2, 3)])
def test_hash_encode_other_feature_widths(F, L):
    emb, tables = make_embedder(cases.BBOX_UNIT, 11, L=L, F=F)
    x = cases.points_in_box(3000, cases.BBOX_UNIT, 31)
    res = O.level_resolutions(16, 512, L)
    want, _ = O.hash_encode(t(x), t(tables), t(np.float32(cases.BBOX_UNIT[0])), t(np.float32(cases.BBOX_UNIT[1])),
                            res, 11)
    out, _ = emb(g32(x))
    bit_equal(out, want)
    dy = np.random.RandomState(6).randn(3000, L * F).astype(np.float32)
    (out * g32(dy)).sum().backward()
    an = O.hash_encode_grad_tables(t(x), t(dy), t(np.float32(cases.BBOX_UNIT[0])), t(np.float32(cases.BBOX_UNIT[1])),
                                   res, 11, F)
    got = torch.stack([e.weight.grad for e in emb.embeddings]).double().cpu()
    assert (got - an).abs().max().item() <= GRAD_RTOL * an.abs().max().item()


@pytest.mark.parametrize("log2T", [19, 14, 22])
def test_hash_encode_full_size_properties(log2T):
    """BASELINE config 2 sizes (2^24 points, T in {14, 19, 22}): properties that need no full-size oracle."""
    n = 1 << 24
    emb, tables = make_embedder(cases.BBOX_UNIT, log2T)
    assert emb.coherent is None  # default policy: this many points take the sorted (coherent) path
    gen = torch.Generator(device=DEV).manual_seed(0)
    x = torch.rand(n, 3, device=DEV, generator=gen) * 3.0 - 1.5
    out, keep = emb(x)
    assert bool(keep.all())
    # (1) a random subset agrees bit-for-bit with the oracle
    idx = torch.randint(0, n, (4096,), device=DEV, generator=gen)
    want, _ = O.hash_encode(x[idx].cpu(), t(tables), t(np.float32(cases.BBOX_UNIT[0])),
                            t(np.float32(cases.BBOX_UNIT[1])), O.level_resolutions(), log2T)
    bit_equal(out[idx], want)
    # (2) linearity in the tables: scaling by a power of two scales every rounding step exactly
    with torch.no_grad():
        for e in emb.embeddings:
            e.weight.mul_(4.0)
    out4, _ = emb(x)
    bit_equal(out4.detach(), (out.detach() * 4.0))
    # (3) backward conserves mass: corner weights sum to 1, so sum(dtable[l]) == sum(dy[:, l])
    dy = torch.ones(n, 32, device=DEV)
    for e in emb.embeddings:
        e.weight.grad = None
    out4.backward(dy)
    sums = torch.stack([e.weight.grad.double().sum() for e in emb.embeddings]).cpu().numpy()
    np.testing.assert_allclose(sums, np.full(16, 2.0 * n), rtol=1e-4)
    # (4) the coherent path and the plain path agree: forward bit-for-bit, gradients to atomic tolerance
    g_sorted = torch.stack([e.weight.grad for e in emb.embeddings]).clone()
    emb.coherent = False
    for e in emb.embeddings:
        e.weight.grad = None
    out_plain, _ = emb(x)
    bit_equal(out_plain.detach(), out4.detach())
    out_plain.backward(dy)
    g_plain = torch.stack([e.weight.grad for e in emb.embeddings])
    assert (g_plain - g_sorted).abs().max().item() <= GRAD_RTOL * g_plain.abs().max().item()


# ---------------------------------------------------------------------------------------------- SH
def test_sh_golden(golden):
    from embedding.spherical_harmonic import SHEncoder
    g = golden("sh")
    for deg in (1, 2, 3, 4, 5):
        bit_equal(SHEncoder(degree=deg)(g32(g["dirs"])), g[f"deg{deg}"])
    out = SHEncoder()(g32(g["dirs"]).reshape(16, 32, 3))
    assert out.shape == (16, 32, 16)


# ---------------------------------------------------------------------------------------------- MLP
@pytest.fixture(params=[(1, 1), (1, 0), (0, 0)], ids=["tcgen05_fused_bwd", "tcgen05_two_kernel_bwd", "ffma_fp32"])
def mlp_impl(request):
    """Every MLP implementation must meet the same bar: tcgen05 forward + the fused bf16x2 backward (default),
    tcgen05 forward + the two-kernel 3xTF32 backward, and the FFMA fp32 one."""
    from hn_b200 import _lib
    _lib.set_tuning("mlp_impl", request.param[0])
    _lib.set_tuning("mlp_bwd_impl", request.param[1])
    yield request.param
    _lib.set_tuning("mlp_impl", 1)
    _lib.set_tuning("mlp_bwd_impl", 1)


def test_mlp_golden(golden, mlp_impl):
    g = golden("mlp")
    net = make_mlp([g[f"w{i}"] for i in range(5)])
    x = g32(g["x"]).requires_grad_(True)
    out = net(x)
    close(out, g["out"], FWD_RTOL, atol=1e-6, what="mlp forward")
    (out * g32(g["dout"])).sum().backward()
    ws = list(net.sigma_net) + list(net.color_net)
    close(x.grad[:, :32], g["dx"][:, :32], GRAD_RTOL, atol=1e-5, what="d_enc")
    for i in range(5):
        close(ws[i].weight.grad, g[f"dw{i}"], GRAD_RTOL, atol=2e-5 * np.abs(g[f"dw{i}"]).max(), what=f"dW{i}")


@pytest.mark.parametrize("n,per_ray", [(1, 1), (127, 1), (129, 1), (64 * 50, 64), (192 * 7, 192)])
def test_mlp_vs_oracle(n, per_ray, mlp_impl):
    sig, col = cases.mlp_weights(70 + n % 7)
    net = make_mlp(sig + col)
    rs = np.random.RandomState(n)
    enc = (rs.randn(n, 32) * 0.3).astype(np.float32)
    views = rs.randn(n // per_ray, 16).astype(np.float32)
    keep = rs.rand(n) > 0.1
    dout = rs.randn(n, 4).astype(np.float32)
    # oracle
    ws = [t(w).requires_grad_(True) for w in sig + col]
    e_t = t(enc).requires_grad_(True)
    full = torch.cat([e_t, t(views).repeat_interleave(per_ray, 0)], -1)
    o = O.nerf_small(full, ws[:2], ws[2:])
    o = torch.cat([o[:, :3], torch.where(t(keep), o[:, 3], torch.zeros(()))[:, None]], -1)
    (o * t(dout)).sum().backward()
    # CUDA
    e_g = g32(enc).requires_grad_(True)
    out = net.forward_fused(e_g, g32(views), per_ray, t(keep).to(DEV))
    close(out, o, FWD_RTOL, atol=1e-6)
    (out * g32(dout)).sum().backward()
    close(e_g.grad, e_t.grad, GRAD_RTOL, atol=1e-5)
    for w_o, lin in zip(ws, list(net.sigma_net) + list(net.color_net)):
        close(lin.weight.grad, w_o.grad, GRAD_RTOL, atol=2e-5 * w_o.grad.abs().max().item())


@pytest.mark.parametrize("n,per_ray", [(128 * 296 * 3 + 77, 1), (192 * 2048, 192), (128 * 5, 64)])
def test_mlp_fused_backward_matches_two_kernel_backward(n, per_ray):
    """The fused bf16x2 backward (one kernel, weight gradients formed on chip) against the 3xTF32 two-kernel backward
    on batches that give every CTA several tiles per context, a ragged last tile and a keep mask."""
    from hn_b200 import _lib
    sig, col = cases.mlp_weights(9)
    rs = np.random.RandomState(n % 1000)
    enc = g32((rs.randn(n, 32) * 0.3).astype(np.float32))
    views = g32(rs.randn((n + per_ray - 1) // per_ray, 16).astype(np.float32))
    keep = torch.from_numpy(rs.rand(n) > 0.1).to(DEV)
    dout = g32(rs.randn(n, 4).astype(np.float32))
    got = {}
    try:
        for impl in (1, 0):
            _lib.set_tuning("mlp_bwd_impl", impl)
            net = make_mlp(sig + col)
            e = enc.clone().requires_grad_(True)
            (net.forward_fused(e, views, per_ray, keep) * dout).sum().backward()
            got[impl] = (e.grad.clone(),
                         torch.cat([l.weight.grad.reshape(-1) for l in list(net.sigma_net) + list(net.color_net)]))
    finally:
        _lib.set_tuning("mlp_bwd_impl", 1)
    for a, b, what in zip(got[1], got[0], ("d_enc", "dW")):
        err = float((a - b).abs().max())
        assert err <= GRAD_RTOL * float(b.abs().max()), (what, err, float(b.abs().max()))
    # elementwise on the weight gradients too: sums over many points average the operand rounding away
    dw1, dw0 = got[1][1], got[0][1]
    assert float(((dw1 - dw0).abs() / (dw0.abs() + 1e-2 * dw0.abs().max())).max()) < 1e-3


def test_nerfsmall_default_geometry_matches_reference_forward():
    """NeRFSmall() with the constructor's OWN defaults (models.py:97-104: 3 sigma layers, 4 colour layers, input_ch 3)
    -- a geometry the fused kernels do not cover -- keeps the reference's behaviour through the layer-by-layer CUDA
    path: parameter names / shapes and outputs equal the layer arithmetic of models.py:151-174."""
    from models import NeRFSmall
    torch.manual_seed(2)
    net = NeRFSmall().to(DEV)
    assert not net.fused
    assert [tuple(l.weight.shape) for l in net.sigma_net] == [(64, 3), (64, 64), (16, 64)]
    assert [tuple(l.weight.shape) for l in net.color_net] == [(64, 18), (64, 64), (64, 64), (3, 64)]
    x = torch.randn(100, 6, device=DEV, requires_grad=True)
    out = net(x)
    w = [l.weight.detach().cpu().double() for l in list(net.sigma_net) + list(net.color_net)]
    xd = x.detach().cpu().double()
    h = torch.relu(torch.relu(xd[:, :3] @ w[0].T) @ w[1].T) @ w[2].T
    sigma, geo = h[:, 0], h[:, 1:]
    c = torch.cat([xd[:, 3:], geo], -1)
    c = torch.relu(torch.relu(torch.relu(c @ w[3].T) @ w[4].T) @ w[5].T) @ w[6].T
    want = torch.cat([c, sigma[:, None]], -1)
    close(out, want.float(), 1e-4, atol=1e-5)
    out.sum().backward()
    assert x.grad is not None and net.sigma_net[0].weight.grad is not None
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NeRFSmall()(torch.randn(4, 6))


def test_mlp_weight_kernel_variants_agree():
    """The two launch shapes of the weight-gradient kernel (mlp_dw_nbuf: two CTAs/SM with one staging buffer, one
    CTA/SM with two) must produce the same gradients, on a batch large enough to keep every CTA busy for a few tiles."""
    from hn_b200 import _lib
    sig, col = cases.mlp_weights(3)
    rs = np.random.RandomState(12)
    n, per_ray = 192 * 400, 192
    enc = g32((rs.randn(n, 32) * 0.3).astype(np.float32))
    views = g32(rs.randn(n // per_ray, 16).astype(np.float32))
    dout = g32(rs.randn(n, 4).astype(np.float32))
    grads = {}
    try:
        _lib.set_tuning("mlp_bwd_impl", 0)
        for nbuf in (1, 2):
            _lib.set_tuning("mlp_dw_nbuf", nbuf)
            net = make_mlp(sig + col)
            (net.forward_fused(enc, views, per_ray, None) * dout).sum().backward()
            grads[nbuf] = torch.cat([l.weight.grad.reshape(-1) for l in list(net.sigma_net) + list(net.color_net)])
    finally:
        _lib.set_tuning("mlp_dw_nbuf", 1)
        _lib.set_tuning("mlp_bwd_impl", 1)
    assert float((grads[1] - grads[2]).abs().max()) <= GRAD_RTOL * float(grads[1].abs().max())


def test_single_pass_sort_for_fine_grids():
    """Grids finer than 256^3 cells use the single-pass counting sort: still a permutation that keeps the coordinates,
    and the sorted kernels produce the plain path's results on it."""
    from hn_b200 import ops
    emb, tables = make_embedder(cases.BBOX_ODD, 12)
    n = 70_001
    x = g32(cases.points_in_box(n, cases.BBOX_ODD, seed=8))
    box, res = emb._geometry(torch.device(DEV))
    xs4 = ops.hash_sort_points(x, box, 300)
    rows = xs4[:, 3].contiguous().view(torch.int32).long()
    assert torch.equal(torch.sort(rows).values, torch.arange(n, device=DEV))
    bit_equal(xs4[:, :3], x[rows])
    flat = emb.flat_tables().reshape(-1)
    want, _ = ops.hash_encode_forward(x, flat, box, res, 16, 2, 12)
    got, _ = ops.hash_encode_forward_sorted(xs4, flat, box, res, 16, 2, 12)
    bit_equal(got, want)


def test_mlp_plain_autograd_gradients_match_sink_gradients():
    """NeRFSmall.fused_grad_accumulation = False returns dense gradients through autograd instead of accumulating into
    the persistent buffer; both must give the same numbers (and accumulate over two backward passes)."""
    sig, col = cases.mlp_weights(4)
    rs = np.random.RandomState(5)
    n, per_ray = 64 * 30, 64
    enc = g32((rs.randn(n, 32) * 0.3).astype(np.float32))
    views = g32(rs.randn(n // per_ray, 16).astype(np.float32))
    dout = g32(rs.randn(n, 4).astype(np.float32))
    got = {}
    for fused in (True, False):
        net = make_mlp(sig + col)
        net.fused_grad_accumulation = fused
        for _ in range(2):
            (net.forward_fused(enc, views, per_ray, None) * dout).sum().backward()
        got[fused] = torch.cat([l.weight.grad.reshape(-1) for l in list(net.sigma_net) + list(net.color_net)]).clone()
    assert float((got[True] - got[False]).abs().max()) <= GRAD_RTOL * float(got[True].abs().max())


def test_render_options_staticcam_and_render_factor():
    """render(c2w_staticcam=...) renders the static camera's rays with the moving camera's view directions
    (run_nerf_helpers.py:351-355); render_path(render_factor=2) renders at half resolution."""
    from embedding.spherical_harmonic import SHEncoder
    from run_nerf_helpers import render, render_path, run_network
    emb, _ = make_embedder(cases.BBOX_UNIT, 12, scale=3000.0)
    sig, col = cases.mlp_weights(6)
    net, sh = make_mlp(sig + col), SHEncoder()
    qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)
    H, W, focal = 8, 10, 12.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    cam = g32(np.array([[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 4.0]], np.float32))
    a = np.deg2rad(25.0)                                  # a ROTATED camera: only rotations change view directions
    moved = g32(np.array([[np.cos(a), 0, np.sin(a), 0.7], [0, 1, 0, 0.0], [-np.sin(a), 0, np.cos(a), 4.0]], np.float32))
    kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=net, network_query_fn=qfn, N_samples=16,
              embed_fn=emb, perturb=0., N_importance=0, white_bkgd=True)
    with torch.no_grad():
        plain, *_ = render(H, W, K, chunk=64, c2w=cam, **kw)
        same_cam, *_ = render(H, W, K, chunk=64, c2w=cam, c2w_staticcam=cam, **kw)
        static, *_ = render(H, W, K, chunk=64, c2w=moved, c2w_staticcam=cam, **kw)
        rgbs, depths = render_path(torch.stack([torch.cat([cam, g32(np.array([[0, 0, 0, 1.0]], np.float32))])]),
                                   (H, W, focal), K, 64, dict(kw, near=2., far=6.), render_factor=2)
    bit_equal(same_cam, plain)
    assert static.shape == plain.shape and bool(torch.isfinite(static).all())
    assert not torch.equal(static, plain)          # view-dependent colour: same geometry, other directions
    assert rgbs.shape == (1, H // 2, W // 2, 3) and depths.shape == (1, H // 2, W // 2)


def test_run_network_generic_path_matches_fused_path():
    """run_network with callables it cannot fuse (here: NeRFSmall hidden behind a plain function, and the SH encoder
    behind a lambda) takes the reference's expand / cat / chunked-apply / mask sequence (run_nerf_helpers.py:212-227)
    and must agree with the fused path, including the zeroed sigma where the keep mask is False."""
    from embedding.spherical_harmonic import SHEncoder
    from run_nerf_helpers import run_network
    emb, _ = make_embedder(cases.BBOX_UNIT, 12, scale=3000.0)
    sig, col = cases.mlp_weights(9)
    net = make_mlp(sig + col)
    sh = SHEncoder()
    rs = np.random.RandomState(2)
    R, S = 37, 24
    pts = rs.uniform(-1.6, 1.6, size=(R, S, 3)).astype(np.float32)
    pts[0, 0] = [np.nan, 0.0, 0.0]                       # the only way a keep mask goes False at L = 16
    dirs = rs.randn(R, 3).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    with torch.no_grad():
        fused = run_network(g32(pts), g32(dirs), net, embed_fn=emb, embeddirs_fn=sh)
        generic = run_network(g32(pts), g32(dirs), lambda x: net(x), embed_fn=emb, embeddirs_fn=lambda d: sh(d),
                              netchunk=200)
    assert fused.shape == generic.shape == (R, S, 4)
    assert float(generic[0, 0, 3]) == 0.0 and float(fused[0, 0, 3]) == 0.0
    ok = torch.isfinite(fused).all(-1)
    close(generic[ok], fused[ok], FWD_RTOL, atol=1e-6)


# ---------------------------------------------------------------------------------------------- compositing
@pytest.mark.parametrize("tag,white", [("black", False), ("white", True)])
def test_composite_golden(golden, tag, white):
    from run_nerf_helpers import raw2outputs
    g = golden("composite")
    raw = g32(g["raw"]).requires_grad_(True)
    rgb, disp, acc, wts, depth, ent = raw2outputs(raw, g32(g["z"]), g32(g["rays_d"]), 0, white)
    for name, val, atol in (("rgb", rgb, 1e-6), ("acc", acc, 1e-6), ("weights", wts, 1e-7), ("depth", depth, 1e-5),
                            ("disp", disp, 1e-6), ("entropy", ent, 2e-6)):
        close(val, g[f"{tag}_{name}"], FWD_RTOL, atol=atol, what=name)
    assert torch.isnan(depth[0])                       # sum(w) == 0 -> NaN depth, like the reference
    good = torch.ones(raw.shape[0], dtype=torch.bool, device=DEV)
    good[0] = False
    good[2] = False
    wm = g32(g["w_misc"])
    loss = (rgb * g32(g["w_rgb"]))[good].sum() + (acc * wm[:, 0])[good].sum() + (depth * wm[:, 1])[good].sum() \
        + (ent * wm[:, 2])[good].sum() + (wts * g32(g["w_wts"]))[good].sum()
    loss.backward()
    want = g[f"{tag}_draw"]
    close(raw.grad, want, GRAD_RTOL, atol=2e-5 * np.abs(want[np.isfinite(want)]).max(), what="d_raw")


def test_composite_noise_and_disp_grad(golden):
    g = golden("composite")
    from hn_b200 import ops
    raw_c = t(g["raw"])[4:].clone().requires_grad_(True)       # rows with sum(w) > 0 only
    z, d, noise = t(g["z"])[4:], t(g["rays_d"])[4:], t(g["noise"])[4:]
    o = O.composite(raw_c, z, d, noise, True)
    w = [torch.randn(v.shape, generator=torch.Generator().manual_seed(i)) for i, v in enumerate(o)]
    sum((a * b).sum() for a, b in zip(o, w)).backward()
    raw_g = raw_c.detach().to(DEV).requires_grad_(True)
    got = ops.CompositeFn.apply(raw_g, z.to(DEV), d.to(DEV), noise.to(DEV), True)
    for a, b in zip(got, o):
        close(a, b, FWD_RTOL, atol=2e-6)
    sum((a * b.to(DEV)).sum() for a, b in zip(got, w)).backward()
    close(raw_g.grad, raw_c.grad, GRAD_RTOL, atol=2e-5 * raw_c.grad.abs().max().item())


@pytest.mark.parametrize("S", [1, 2, 31, 33, 64, 192, 257])
def test_composite_sample_counts(S):
    from hn_b200 import ops
    rs = np.random.RandomState(S)
    R = 37
    raw = rs.randn(R, S, 4).astype(np.float32)
    raw[..., 3] = raw[..., 3] * 2 + 1
    z = np.sort(2 + 4 * rs.rand(R, S).astype(np.float32), -1)
    d = rs.randn(R, 3).astype(np.float32)
    want = O.composite(t(raw), t(z), t(d), None, False)
    got = ops.CompositeFn.apply(g32(raw), g32(z), g32(d), None, False)
    for a, b in zip(got, want):
        close(a, b, FWD_RTOL, atol=2e-6)


# ---------------------------------------------------------------------------------------------- sampling
def test_sample_pdf_golden(golden):
    from hn_b200 import ops
    g = golden("sample_pdf")
    got = ops.sample_pdf(g32(g["bins"]), g32(g["weights"]), g["u"].shape[1], u=g32(g["u"]))
    want, _, _ = O.sample_pdf_tolerance(t(g["bins"]), t(g["weights"]), t(g["u"]))
    bit_equal(want, g["samples_rand"])          # the bound is centred on the reference's own output
    samples_within_bound(got, g["bins"], g["weights"], g["u"], what="random u")
    from run_nerf_helpers import sample_pdf
    got = sample_pdf(g32(g["bins"]), g32(g["weights"]), g["u"].shape[1], det=True)
    u_det = O.det_u(g["bins"].shape[0], g["u"].shape[1]).contiguous().numpy()
    bit_equal(O.sample_pdf(t(g["bins"]), t(g["weights"]), t(u_det)), g["samples_det"])
    samples_within_bound(got, g["bins"], g["weights"], u_det, what="det")


@pytest.mark.parametrize("R,S,Ni,det", [(50, 64, 128, False), (33, 64, 64, True), (7, 24, 40, False), (5, 3, 1, False)])
def test_resample_fused(R, S, Ni, det):
    """The fused resampling launch == mids + sample_pdf(weights[:,1:-1]) + sort(cat) + std, stage by stage."""
    from hn_b200 import ops
    rs = np.random.RandomState(R + S)
    z = np.sort(2 + 4 * rs.rand(R, S).astype(np.float32), -1)
    w = (rs.rand(R, S).astype(np.float32)) ** 4
    u = None if det else rs.rand(R, Ni).astype(np.float32)
    mids = 0.5 * (t(z)[:, 1:] + t(z)[:, :-1])
    want = O.sample_pdf(mids, t(w)[:, 1:-1], O.det_u(R, Ni) if det else t(u))
    kw = dict(u_det=torch.linspace(0., 1., steps=Ni, device=DEV)) if det else dict(u=g32(u))
    samples, merged, z_std = ops.resample(g32(z), g32(w), Ni, **kw)
    u_used = O.det_u(R, Ni).contiguous() if det else t(u)
    samples_within_bound(samples, mids.numpy(), w[:, 1:-1], u_used.numpy(), what="samples")
    bit_equal(merged, torch.sort(torch.cat([g32(z), samples], -1), -1).values)
    close(z_std, torch.std(samples, dim=-1, unbiased=False), 1e-4, atol=1e-6)


def test_resample_unsorted_depths_and_duplicates():
    """The merge-by-rank path needs sorted coarse depths; rows that are not (or that repeat values between the
    depths and the new samples) must still come out as sort(cat(z, samples))."""
    from hn_b200 import ops
    rs = np.random.RandomState(11)
    R, S, Ni = 37, 24, 40
    z = np.sort(2 + 4 * rs.rand(R, S).astype(np.float32), -1)
    z[::3] = z[::3, ::-1]                       # every third row descending: takes the full network
    z[1, 5] = z[1, 4]                           # a tie inside the depths
    w = (rs.rand(R, S).astype(np.float32)) ** 2
    w[2, :] = 0.0                               # all-empty row: every denominator below 1e-5
    u = rs.rand(R, Ni).astype(np.float32)
    u[4, :] = 0.5                               # Ni identical samples
    samples, merged, _ = ops.resample(g32(z), g32(w), Ni, u=g32(u))
    bit_equal(merged, torch.sort(torch.cat([g32(z), samples], -1), -1).values)


def test_img2mse_one_launch():
    """img2mse on CUDA goes through hn_mse_fwd / hn_mse_bwd: value and both gradients against the reference's
    expression (run_nerf_helpers.py:24)."""
    from run_nerf_helpers import img2mse
    for shape in [(1024, 3), (7,), (5000, 3)]:
        gen = torch.Generator(device=DEV).manual_seed(len(shape))
        a = torch.rand(*shape, device=DEV, generator=gen).requires_grad_(True)
        b = torch.rand(*shape, device=DEV, generator=gen).requires_grad_(True)
        got = img2mse(a, b)
        (3.0 * got).backward()
        ga, gb = a.grad.clone(), b.grad.clone()
        a.grad = b.grad = None
        want = torch.mean((a - b) ** 2)
        (3.0 * want).backward()
        close(got, want, 1e-6)
        close(ga, a.grad, 1e-6, atol=1e-9)
        close(gb, b.grad, 1e-6, atol=1e-9)
    # CPU tensors keep the reference's expression
    x, y = torch.rand(4, 3), torch.rand(4, 3)
    assert torch.equal(img2mse(x, y), torch.mean((x - y) ** 2))


def test_sort_concat_rows():
    from hn_b200 import ops
    for na, nb, R in [(64, 128, 100), (1, 1, 3), (5, 0, 4), (24, 40, 48), (700, 1300, 5), (64, 64, 1000)]:
        gen = torch.Generator(device=DEV).manual_seed(na + nb)
        a = torch.rand(R, na, device=DEV, generator=gen).sort(-1).values
        b = torch.rand(R, nb, device=DEV, generator=gen)
        if na > 3 and nb > 3:
            b[:, 0] = a[:, 2]  # duplicates
        want = torch.sort(torch.cat([a, b], -1), -1).values
        bit_equal(ops.sort_concat_rows(a, b), want)


@pytest.mark.parametrize("lindisp,perturb", [(False, False), (False, True), (True, True)])
def test_coarse_z_and_points_bit_exact(lindisp, perturb):
    from hn_b200 import ops
    R, S = 77, 64
    rays = cases.rays(R, 5)
    rays[:, 6] = 0.5 + np.random.RandomState(1).rand(R).astype(np.float32)
    t_rand = np.random.RandomState(2).rand(R, S).astype(np.float32) if perturb else None
    rb = g32(rays)
    tv = torch.linspace(0., 1., steps=S, device=DEV)
    bit_equal(tv, torch.linspace(0., 1., steps=S))
    z = ops.coarse_z(rb[:, 6], rb[:, 7], 11, tv, None if t_rand is None else g32(t_rand), R, S, lindisp)
    want = O.coarse_z(t(rays[:, 6:7]), t(rays[:, 7:8]), S, lindisp, None if t_rand is None else t(t_rand))
    bit_equal(z, want.contiguous())
    pts = ops.ray_points(rb[:, 0:3], rb[:, 3:6], 11, z)
    want_pts = t(rays[:, None, 0:3]) + t(rays[:, None, 3:6]) * want[:, :, None]
    bit_equal(pts, want_pts)


# ---------------------------------------------------------------------------------------------- end to end
@pytest.mark.parametrize("name", ["render_rays_perturb", "render_rays_det_noise"])
def test_render_rays_golden(golden, name, monkeypatch):
    from embedding.spherical_harmonic import SHEncoder
    from run_nerf_helpers import render_rays, run_network
    g = golden(name)
    emb, _ = make_embedder(g["bbox"], int(g["log2T"]), scale=float(g["table_scale"]))
    coarse = make_mlp([g[f"coarse_w{i}"] for i in range(5)])
    fine = make_mlp([g[f"fine_w{i}"] for i in range(5)])
    sh = SHEncoder()
    qfn = lambda inputs, viewdirs, fn: run_network(inputs, viewdirs, fn, embed_fn=emb, embeddirs_fn=sh)
    import run_nerf_helpers as H
    from hn_b200 import ops
    kw = dict(embed_fn=emb, retraw=True, perturb=float(g["perturb"]), N_importance=int(g["N_importance"]),
              network_fine=fine, white_bkgd=bool(g["white_bkgd"]), raw_noise_std=float(g["raw_noise_std"]), pytest=True)
    real_resample, spy = ops.resample, {}

    def recording(z_vals, weights, n, u=None, u_det=None):
        out = real_resample(z_vals, weights, n, u=u, u_det=u_det)
        spy["call"] = (z_vals.clone(), weights.clone(), n, u, u_det, out[0].clone())
        return out

    def injecting(z_vals, weights, n, u=None, u_det=None):
        zs = g32(g["z_samples"])
        return zs, ops.sort_concat_rows(z_vals, zs), torch.std(zs, dim=-1, unbiased=False)

    with torch.no_grad():
        monkeypatch.setattr(H.ops, "resample", recording)
        render_rays(g32(g["rays"]), coarse, qfn, int(g["N_samples"]), **kw)       # our own resampling
    monkeypatch.setattr(H.ops, "resample", injecting)
    ret = ret_inj = render_rays(g32(g["rays"]), coarse, qfn, int(g["N_samples"]), **kw)   # the reference's sample positions
    monkeypatch.setattr(H.ops, "resample", real_resample)
    # coarse pass: no resampling involved -> strict tolerance
    for k in ("rgb0", "depth0", "acc0", "sparsity_loss0"):
        want = g["ret_" + k]
        close(ret[k], want, 5e-5, atol=2e-5 * max(1.0, float(np.abs(want).max())), what=k)
    # the resampled depths: the oracle restates the reference's z_samples bit for bit on the reference's inputs ...
    R, Ni = g["z_samples"].shape
    np.random.seed(0)
    u = (np.broadcast_to(np.linspace(0., 1., Ni), (R, Ni)) if float(g["perturb"]) == 0. else np.random.rand(R, Ni))
    u = np.ascontiguousarray(u).astype(np.float32)
    want_z, _, _ = O.sample_pdf_tolerance(t(g["pdf_bins"]), t(g["pdf_weights"]), t(u))
    bit_equal(want_z, g["z_samples"])
    # ... the variates we drew are the reference's, and every depth we resampled lies inside its analytic bound
    # around that oracle evaluated on the inputs OUR coarse pass produced
    bit_equal(spy["call"][3] if spy["call"][3] is not None else spy["call"][4].expand(R, Ni), u)
    recorded_resample_within_bound(spy["call"], what="z_samples")
    # fine pass on the REFERENCE's sample positions: strict tolerance on everything downstream
    for k in ("rgb_map", "depth_map", "acc_map", "sparsity_loss", "z_std", "raw"):
        want = g["ret_" + k]
        close(ret_inj[k], want, 5e-5, atol=2e-5 * max(1.0, float(np.abs(want).max())), what=k + " (reference z_samples)")
    tgt = g32(g["target"])
    loss = ((ret_inj["rgb_map"] - tgt) ** 2).mean() + ((ret_inj["rgb0"] - tgt) ** 2).mean() \
        + 1e-3 * (ret_inj["sparsity_loss"].sum() + ret_inj["sparsity_loss0"].sum())
    close(loss, g["loss"], 1e-5)
    loss.backward()
    gt = torch.stack([e.weight.grad for e in emb.embeddings])
    want = g["grad_tables"]
    close(gt, want, GRAD_RTOL, atol=GRAD_RTOL * np.abs(want).max(), what="table grads")
    for tag, net in (("coarse", coarse), ("fine", fine)):
        for i, lin in enumerate(list(net.sigma_net) + list(net.color_net)):
            want = g[f"{tag}_dw{i}"]
            close(lin.weight.grad, want, GRAD_RTOL, atol=GRAD_RTOL * np.abs(want).max(), what=f"{tag} dW{i}")


# ---------------------------------------------------------------------------------------------- rays / render
def _pose(seed):
    rs = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rs.randn(3, 3))
    return np.concatenate([q, rs.randn(3, 1) * 2], -1).astype(np.float32)


def test_get_rays_and_pack_rays():
    from ray_util import get_rays
    from hn_b200 import ops
    H, W, focal = 37, 53, 61.5
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = _pose(3)
    o_cpu, d_cpu = get_rays(H, W, K, t(c2w))                  # plain torch branch (CPU tensor)
    o_gpu, d_gpu = get_rays(H, W, K, g32(c2w))                # hn_get_rays
    assert d_gpu.shape == (H, W, 3) and o_gpu.shape == (H, W, 3)
    close(d_gpu, d_cpu, 1e-6, atol=1e-6)
    bit_equal(o_gpu.contiguous(), o_cpu.contiguous())
    rays_o, rays_d = o_gpu.reshape(-1, 3), d_gpu.reshape(-1, 3)
    packed = ops.pack_rays(rays_o, rays_d, rays_d, 2.0, 6.0)
    assert packed.shape == (H * W, 11)
    bit_equal(packed[:, 0:3].contiguous(), rays_o.contiguous())
    bit_equal(packed[:, 3:6].contiguous(), rays_d.contiguous())
    assert bool((packed[:, 6] == 2.0).all()) and bool((packed[:, 7] == 6.0).all())
    close(packed[:, 8:], rays_d / torch.norm(rays_d, dim=-1, keepdim=True), 1e-6, atol=1e-7)
    assert ops.pack_rays(rays_o, rays_d, None, 0.0, 1.0).shape == (H * W, 8)


def test_get_rays_golden(golden):
    from ray_util import get_rays, get_ndc_rays
    g = golden("rays")
    H, W, focal = int(g["H"]), int(g["W"]), float(g["focal"])
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    o, d = get_rays(H, W, K, g32(g["c2w"]))                       # hn_get_rays
    bit_equal(d, g["rays_d"])                                      # same roundings as the reference's chain
    bit_equal(o.contiguous(), g["rays_o"])
    of, df = get_rays(H, W, K, g32(g["c2w_f"]))
    no, nd = get_ndc_rays(H, W, K[0][0], 1., of.reshape(-1, 3), df.reshape(-1, 3))   # hn_ndc_rays
    bit_equal(no, g["ndc_o"])
    bit_equal(nd, g["ndc_d"])
    no2, nd2 = get_ndc_rays(H, W, K[0][0], 1., of, df)                               # [H,W,3] in, [H,W,3] out
    bit_equal(no2.reshape(-1, 3), g["ndc_o"])
    bit_equal(nd2.reshape(-1, 3), g["ndc_d"])


@pytest.mark.parametrize("n_rand,crop", [(1024, False), (300, True), (37 * 53, False)])
def test_sample_rays_device_batcher(n_rand, crop):
    """hn_sample_rays (opt-in replacement of run_nerf.py:576-605 + run_nerf_helpers.py:344-366): n_rand DISTINCT
    pixels inside the window; the packed rays are bit-identical to get_rays + pack_rays at those pixels and the
    targets are the image values there."""
    from ray_util import get_rays
    from hn_b200 import ops
    from hn_b200.batcher import DeviceRayBatcher
    H, W, focal, n_img = 37, 53, 61.5, 5
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    rs = np.random.RandomState(9)
    images = rs.rand(n_img, H, W, 3).astype(np.float32)
    poses = np.stack([np.concatenate([_pose(10 + i), np.array([[0, 0, 0, 1]], np.float32)], 0) for i in range(n_img)])
    b = DeviceRayBatcher(images, poses, H, W, K, 2.0, 6.0, n_rand, DEV, i_train=[1, 3, 4],
                         precrop_iters=5 if crop else 0, precrop_frac=0.8, seed=4)
    seen_images, first = set(), None
    for _ in range(3):
        rays, target, pix = b.next(want_pix=True)
        torch.cuda.synchronize()
        img = int(b._dev[0].item())
        seen_images.add(img)
        assert img in (1, 3, 4)
        row0, col0, wh, ww = b.window(b.step_index - 1)
        p = pix.cpu().numpy().astype(np.int64)
        assert len({(int(r), int(c)) for r, c in p}) == n_rand, "pixels must be distinct (sampling without replacement)"
        assert p[:, 0].min() >= row0 and p[:, 0].max() < row0 + wh and p[:, 1].min() >= col0 and p[:, 1].max() < col0 + ww
        o_all, d_all = get_rays(H, W, K, g32(poses[img, :3, :4]))
        sel_d = d_all[p[:, 0], p[:, 1]]
        sel_o = o_all[p[:, 0], p[:, 1]]
        want = ops.pack_rays(sel_o.contiguous(), sel_d.contiguous(), sel_d.contiguous(), 2.0, 6.0)
        bit_equal(rays, want)
        bit_equal(target, images[img][p[:, 0], p[:, 1]])
        if first is None:
            first = p.copy()
        else:
            assert not np.array_equal(first, p), "a new permutation every step"
    if n_rand == H * W:   # a full permutation of the image
        assert sorted(map(tuple, p.tolist())) == [(r, c) for r in range(H) for c in range(W)]


def test_render_ndc_forward_facing_against_oracle(monkeypatch):
    """LLFF-style path (BASELINE configs[4], fern.txt shapes scaled down): NDC warp, N_importance = N_samples,
    no white background; every output strict, the resampled depths inside their analytic bound."""
    from embedding.spherical_harmonic import SHEncoder
    from run_nerf_helpers import render, run_network
    H, W, focal = 9, 12, 11.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = np.array([[1, 0, 0, 0.05], [0, 1, 0, -0.02], [0, 0, 1, 0.3]], np.float32)
    bbox = ((-1.6, -1.6, -1.1), (1.6, 1.6, 1.1))
    emb, tables = make_embedder(bbox, 12, scale=3000.0)
    w_c, w_f = sum(cases.mlp_weights(5), []), sum(cases.mlp_weights(6), [])
    coarse, fine, sh = make_mlp(w_c), make_mlp(w_f), SHEncoder()
    qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)
    import run_nerf_helpers as Hm
    from hn_b200 import ops
    real_resample, drawn = ops.resample, []

    def recording(z_vals, weights, n, u=None, u_det=None):
        out = real_resample(z_vals, weights, n, u=u, u_det=u_det)
        drawn.append((z_vals.clone(), weights.clone(), n, u, u_det, out[0].clone()))
        return out

    monkeypatch.setattr(Hm.ops, "resample", recording)
    with torch.no_grad():
        rgb, depth, acc, extras = render(H, W, K, chunk=64, c2w=g32(c2w), ndc=True, near=0., far=1.,
                                         use_viewdirs=True, network_fn=coarse, network_fine=fine,
                                         network_query_fn=qfn, N_samples=16, N_importance=16, embed_fn=emb,
                                         perturb=0., raw_noise_std=0., white_bkgd=False)
    monkeypatch.setattr(Hm.ops, "resample", real_resample)
    for call in drawn:  # stage 1: every chunk's resampled depths inside their bound (oracle on the same inputs)
        recorded_resample_within_bound(call, what="z_samples")
    ours_z = torch.cat([c[5] for c in drawn], 0).cpu()
    o, d = O.pinhole_rays(H, W, K, t(c2w))
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    vd = d / torch.norm(d, dim=-1, keepdim=True)
    no, nd = O.ndc_rays(H, W, K[0][0], 1., o, d)
    rays = torch.cat([no, nd, torch.zeros_like(nd[:, :1]), torch.ones_like(nd[:, :1]), vd], -1).contiguous()
    lo, hi = t(np.float32(bbox[0])), t(np.float32(bbox[1]))
    enc = lambda p: O.hash_encode(p, t(tables), lo, hi, O.level_resolutions(), 12)
    cw, fw = [t(w) for w in w_c], [t(w) for w in w_f]
    # stage 2: the oracle's fine pass on OUR sample positions -> strict tolerance on everything downstream
    want = O.render_rays(rays, enc, (cw[:2], cw[2:]), (fw[:2], fw[2:]), 16, 16, white_bkgd=False, perturb=0.,
                         z_samples=ours_z)
    close(extras["rgb0"].reshape(-1, 3), want["rgb0"], 5e-5, atol=2e-5)
    close(extras["acc0"].reshape(-1), want["acc0"], 5e-5, atol=2e-5)
    close(rgb.reshape(-1, 3), want["rgb_map"], 5e-5, atol=2e-5, what="fine rgb")
    close(acc.reshape(-1), want["acc_map"], 5e-5, atol=2e-5, what="fine acc")


def test_render_full_image_against_oracle():
    """render(c2w=...) -> get_rays -> pack -> chunked render_rays, coarse only, vs the oracle on the same rays."""
    from embedding.spherical_harmonic import SHEncoder
    from run_nerf_helpers import render, run_network
    H, W, focal = 12, 10, 14.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = np.array([[1, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 4.0]], np.float32)
    emb, tables = make_embedder(cases.BBOX_UNIT, 12, scale=3000.0)
    sig, col = cases.mlp_weights(5)
    net = make_mlp(sig + col)
    sh = SHEncoder()
    qfn = lambda i, v, fn: run_network(i, v, fn, embed_fn=emb, embeddirs_fn=sh)
    with torch.no_grad():
        rgb, depth, acc, extras = render(H, W, K, chunk=50, c2w=g32(c2w), ndc=False, near=2., far=6.,
                                         use_viewdirs=True, network_fn=net, network_query_fn=qfn, N_samples=16,
                                         embed_fn=emb, perturb=0., N_importance=0, white_bkgd=True)
    assert rgb.shape == (H, W, 3) and depth.shape == (H, W) and "sparsity_loss" in extras
    from ray_util import get_rays
    o, d = get_rays(H, W, K, t(c2w))
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    rays = torch.cat([o, d, torch.full_like(d[:, :1], 2.), torch.full_like(d[:, :1], 6.), d / d.norm(dim=-1, keepdim=True)], -1)
    lo, hi = t(np.float32(cases.BBOX_UNIT[0])), t(np.float32(cases.BBOX_UNIT[1]))
    enc = lambda p: O.hash_encode(p, t(tables), lo, hi, O.level_resolutions(), 12)
    ws = [t(w) for w in sig + col]
    want = O.render_rays(rays.contiguous(), enc, (ws[:2], ws[2:]), None, 16, 0, white_bkgd=True, perturb=0.)
    close(rgb.reshape(-1, 3), want["rgb_map"], 5e-5, atol=2e-5)
    close(acc.reshape(-1), want["acc_map"], 5e-5, atol=2e-5)


# ---------------------------------------------------------------------------------------------- next rows
def test_radam_fused_zero_grad_and_plan_invalidation():
    """(1) fused_zero_grad=True: same parameters as the plain optimizer, gradients read zero after step(), and the
    hash-table gradient sink skips its own fill on the next backward yet accumulates correctly.  (2) load_state_dict /
    add_param_group drop the cached span plan: a later lr change reaches the kernel and moments that are no longer
    one allocation are not run over as one span (radam.py:66)."""
    from radam import RAdam
    torch.manual_seed(0)
    emb_a, _ = make_embedder(cases.BBOX_UNIT, 10, scale=3000.0)
    emb_b, _ = make_embedder(cases.BBOX_UNIT, 10, scale=3000.0)
    opt_a = RAdam([{"params": list(emb_a.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
    opt_b = RAdam([{"params": list(emb_b.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99), fused_zero_grad=True)
    x = g32(cases.points_in_box(3000, cases.BBOX_UNIT, seed=3))
    for step in range(7):   # crosses RAdam's rectification switch at step 6
        dy = torch.randn(3000, 32, device=DEV, generator=torch.Generator(device=DEV).manual_seed(step))
        for emb, opt in ((emb_a, opt_a), (emb_b, opt_b)):
            opt.zero_grad()
            emb(x)[0].backward(dy)
            opt.step()
        assert float(emb_b.grad_sink().flat.abs().max()) == 0.0 and emb_b.grad_sink().clean
        assert float(emb_a.grad_sink().flat.abs().max()) > 0.0
        # same trajectory (not bit for bit: the scatter's atomic accumulation order differs from run to run)
        close(emb_b.flat_tables(), emb_a.flat_tables(), 1e-5, atol=1e-7)
    # (2) resume from a state dict whose moments are SEPARATE allocations (what a reference checkpoint holds)
    sd = opt_a.state_dict()
    sd = {"state": {k: {kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} for k, v in sd["state"].items()},
          "param_groups": sd["param_groups"]}
    opt_a.load_state_dict(sd)
    assert "_span_cache" not in opt_a.__dict__
    for grp in opt_a.param_groups:
        grp["lr"] = 0.0            # must reach the kernel: with a stale plan the old group dict (lr 0.01) would be used
    before = emb_a.flat_tables().clone()
    opt_a.zero_grad()
    emb_a(x)[0].backward(torch.ones(3000, 32, device=DEV))
    opt_a.step()
    torch.cuda.synchronize()
    bit_equal(emb_a.flat_tables(), before)
    m = [opt_a.state[p]["exp_avg"] for p in emb_a.parameters()]
    assert all(bool(torch.isfinite(t_).all()) for t_ in m)
    opt_a.add_param_group({"params": [torch.nn.Parameter(torch.zeros(4, device=DEV))]})
    assert "_span_cache" not in opt_a.__dict__


def test_radam_golden(golden):
    from radam import RAdam
    g = golden("radam")
    buf = torch.zeros(257 * 3 + 64 * 2, device=DEV)
    p = torch.nn.Parameter(g32(g["p0"]))
    q = torch.nn.Parameter(g32(g["q0"]))
    opt = RAdam([{"params": [p], "weight_decay": 1e-6}, {"params": [q], "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
    for step in range(g["grads_p"].shape[0]):
        p.grad = g32(g["grads_p"][step])
        q.grad = g32(g["grads_q"][step])
        opt.step()
        for grp in opt.param_groups:
            grp["lr"] = 0.01 * (0.1 ** ((step + 1) / 10000.0))
        close(p, g["traj_p"][step], 1e-5, atol=1e-8, what=f"p step {step}")
        close(q, g["traj_q"][step], 1e-5, atol=2e-9, what=f"q step {step}")
    assert set(opt.state[p].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_tv_loss_golden(golden, monkeypatch):
    import loss
    from loss import total_variation_loss
    monkeypatch.setattr(loss, "TV_SWEEP", False)  # per-level path: the cube origin comes from torch.randint
    g = golden("tv_loss")
    emb, tables = make_embedder(cases.BBOX_UNIT, 12)
    real = torch.randint
    for level in (0, 3, 7, 15):
        torch.randint = lambda *a, **k: t(g[f"l{level}_min_vertex"]).clone()  # CPU tensor: loss.py moves it
        try:
            tv = total_variation_loss(emb.embeddings[level], emb.base_resolution, emb.finest_resolution, level, 12,
                                      n_levels=16)
        finally:
            torch.randint = real
        close(tv, g[f"l{level}_tv"], 2e-5)
        emb.embeddings[level].weight.grad = None
        tv.backward()
        dense = np.zeros((1 << 12, 2), np.float32)
        dense[g[f"l{level}_grad_rows"]] = g[f"l{level}_grad_vals"]
        close(emb.embeddings[level].weight.grad, dense, GRAD_RTOL, atol=1e-9)


def test_tv_loss_sweep_matches_per_level_and_caches():
    """All 16 levels in one launch (what the training loop's 16 consecutive calls get) == the per-level kernels on
    the same cubes; and the sweep cache: one evaluation per increasing run of levels, a new one when the run
    restarts, when grad mode changes or after an optimizer step."""
    import loss
    from hn_b200 import _lib, ops
    from radam import RAdam
    emb, tables = make_embedder(cases.BBOX_UNIT, 12)
    L = 16
    geo = [loss._level_cube(emb.base_resolution, emb.finest_resolution, l, L) for l in range(L)]
    rs = np.random.RandomState(3)
    origins = torch.tensor([[rs.randint(0, r - c) for _ in range(3)] for r, c in geo], dtype=torch.int64, device=DEV)
    cubes = torch.tensor([c for _, c in geo], dtype=torch.int32, device=DEV)
    flat = emb.flat_tables()
    vec = ops.TVSweepFn.apply(flat, origins, cubes, max(c for _, c in geo), 12, 2, None, *emb._level_weights())
    gout = g32(rs.rand(L).astype(np.float32))
    (vec * gout).sum().backward()
    got = torch.stack([e.weight.grad for e in emb.embeddings]).clone()
    for e in emb.embeddings:
        e.weight.grad = None
    for l in range(L):
        one = ops.TVLossFn.apply(emb.embeddings[l].weight, origins[l].contiguous(), geo[l][1], 12, None)
        close(vec[l], one, 1e-6)
        (one * gout[l]).backward()
    want = torch.stack([e.weight.grad for e in emb.embeddings])
    close(got, want, GRAD_RTOL, atol=1e-9)
    # the L-scalars form (what the per-level calls are served from): same values, same gradients, also when only
    # some of the terms take part in the loss
    for used in (range(L), (0, 3, 15)):
        for e in emb.embeddings:
            e.weight.grad = None
        parts = ops.TVSweepPartsFn.apply(flat, origins, cubes, max(c for _, c in geo), 12, 2, None, *emb._level_weights())
        assert len(parts) == L and all(p.dim() == 0 for p in parts)
        close(torch.stack(parts), vec.detach(), 1e-5)   # the forward sums with atomics: not bit-reproducible
        sum(parts[l] * gout[l] for l in used).backward()
        got_p = torch.stack([e.weight.grad if e.weight.grad is not None else torch.zeros_like(e.weight)
                             for e in emb.embeddings])
        mask = torch.zeros(L, device=DEV)
        mask[list(used)] = 1.0
        close(got_p, want * mask[:, None, None], GRAD_RTOL, atol=1e-9)
    for e in emb.embeddings:
        e.weight.grad = None

    def sweep_calls(levels, grad=True):
        before = _lib.launches
        with torch.set_grad_enabled(grad):
            terms = [loss.total_variation_loss(emb.embeddings[i], emb.base_resolution, emb.finest_resolution, i, 12,
                                               n_levels=L) for i in levels]
        return terms, _lib.launches - before
    terms, n = sweep_calls(range(L))
    assert n == 1, f"16 consecutive levels must be one launch, saw {n}"
    assert all(tm.requires_grad for tm in terms)
    for e in emb.embeddings:
        e.weight.grad = None
    sum(terms).backward()
    assert all(e.weight.grad is not None and float(e.weight.grad.abs().sum()) > 0 for e in emb.embeddings)
    assert sweep_calls([0, 1, 2])[1] == 1 and sweep_calls([3, 4])[1] == 0      # continues the same sweep
    assert sweep_calls([4])[1] == 1                                             # level repeated: new cubes
    assert sweep_calls([5], grad=False)[1] == 1                                 # grad mode changed
    assert sweep_calls([6, 7], grad=False)[1] == 0
    opt = RAdam(list(emb.parameters()), lr=1e-3)
    sum(sweep_calls([8])[0]).backward()
    opt.step()
    assert sweep_calls([9])[1] == 1, "tables changed (RAdam kernel): the sweep must be re-evaluated"
    with torch.no_grad():
        emb.embeddings[2].weight.mul_(1.0)                                      # torch in-place op: version bump
    assert sweep_calls([10])[1] == 0, "another level's table changed: level 10's term is still the sweep's"
    with torch.no_grad():
        emb.embeddings[11].weight.mul_(1.0)
    assert sweep_calls([11])[1] == 1, "the level's own table changed: the sweep must be re-evaluated"


def test_grad_sink_accumulation_semantics():
    """Backward accumulates in place into one flat buffer: two passes sum, zero_grad in both flavours resets,
    gradients arriving through plain autograd (TV loss via nn.Embedding) are merged, RAdam sees one span."""
    from hn_b200 import _lib
    from radam import RAdam
    emb, tables = make_embedder(cases.BBOX_ODD, 10)
    x1 = g32(cases.points_in_box(3000, cases.BBOX_ODD, 1))
    x2 = g32(cases.points_in_box(5000, cases.BBOX_ODD, 2))
    dy1, dy2 = torch.randn(3000, 32, device=DEV), torch.randn(5000, 32, device=DEV)
    ATOL = 1e-4  # atomically accumulated sums of O(10) values

    def reference(parts):
        ref, _ = make_embedder(cases.BBOX_ODD, 10)
        ref.fused_grad_accumulation = False
        tot = None
        for xx, dd in parts:
            for e in ref.embeddings:
                e.weight.grad = None
            ref(xx)[0].backward(dd)
            g = torch.stack([e.weight.grad for e in ref.embeddings])
            tot = g if tot is None else tot + g
        return tot

    # coarse + fine style: two forward/backward pairs before the optimizer
    (emb(x1)[0] * dy1).sum().backward()
    (emb(x2)[0] * dy2).sum().backward()
    got = torch.stack([e.weight.grad for e in emb.embeddings])
    want = reference([(x1, dy1), (x2, dy2)])
    close(got, want, GRAD_RTOL, atol=1e-6 * want.abs().max().item())
    flat = emb._sink.flat
    assert all(e.weight.grad.data_ptr() == flat[i * 2048:].data_ptr() for i, e in enumerate(emb.embeddings))

    opt = RAdam([{"params": list(emb.parameters()), "eps": 1e-15}], lr=0.01, betas=(0.9, 0.99))
    before = _lib.launches
    opt.step()
    assert _lib.launches - before == 1, "the 16 level tables must be updated by one fused launch"

    opt.zero_grad()                                   # set_to_none=True
    assert all(e.weight.grad is None for e in emb.embeddings)
    (emb(x1)[0] * dy1).sum().backward()
    close(torch.stack([e.weight.grad for e in emb.embeddings]), reference([(x1, dy1)]), GRAD_RTOL, atol=ATOL)

    opt.zero_grad(set_to_none=False)                  # in-place zeroing of our slices
    (emb(x2)[0] * dy2).sum().backward()
    close(torch.stack([e.weight.grad for e in emb.embeddings]), reference([(x2, dy2)]), GRAD_RTOL, atol=ATOL)

    # a gradient that arrives through nn.Embedding first (as the TV loss does) is merged, not lost
    opt.zero_grad()
    idx = torch.arange(0, 64, device=DEV)
    emb.embeddings[3](idx).sum().backward()
    (emb(x1)[0] * dy1).sum().backward()
    want = reference([(x1, dy1)])
    want[3, :64] += 1.0
    close(torch.stack([e.weight.grad for e in emb.embeddings]), want, GRAD_RTOL, atol=ATOL)
    emb.embeddings[3](idx).sum().backward()           # ... and after: autograd adds in place into our slice
    want[3, :64] += 1.0
    close(torch.stack([e.weight.grad for e in emb.embeddings]), want, GRAD_RTOL, atol=ATOL)
    assert emb.embeddings[3].weight.grad.data_ptr() == emb._sink.flat[3 * 2048:].data_ptr()


# ---------------------------------------------------------------------------------------------- errors
def test_error_behaviour():
    from embedding.hash_encoding import HashEmbedder, hash
    from hn_b200 import _lib
    emb, _ = make_embedder(cases.BBOX_UNIT, 10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        emb(torch.rand(4, 3))
    with pytest.raises(RuntimeError, match="dim must be in"):
        hash(torch.zeros(3, 8, dtype=torch.int64, device=DEV), 19)
    with pytest.raises(RuntimeError):
        _lib.set_tuning("no_such_knob", 1)
    with pytest.raises(NotImplementedError):
        from models import NeRF
        NeRF()  # the positional-encoding network of the dead i_embed = 0 branch
