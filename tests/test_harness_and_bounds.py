"""CPU tests of test infrastructure added in round 2: the stand-in modules of the run_nerf.py harness, the scene
generator, and the per-sample acceptance bound of the resampling stage."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STANDINS = os.path.join(ROOT, "tests", "harness", "standins")


def _standin(name):
    import importlib.util
    path = os.path.join(STANDINS, name + ".py")
    if not os.path.isfile(path):
        path = os.path.join(STANDINS, name, "__init__.py")
    spec = importlib.util.spec_from_file_location("_standin_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_png_roundtrip_and_scene(tmp_path):
    imageio = _standin("imageio")
    rs = np.random.RandomState(0)
    for shape in ((7, 5, 4), (6, 9, 3), (4, 4)):
        img = rs.randint(0, 256, size=shape).astype(np.uint8)
        path = os.path.join(tmp_path, "x.png")
        imageio.imwrite(path, img)
        assert np.array_equal(imageio.imread(path), img)
    sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
    import make_scene
    root = make_scene.make(os.path.join(tmp_path, "scene"), H=16, W=16, n_train=3, n_val=1, n_test=1)
    import json
    meta = json.load(open(os.path.join(root, "transforms_train.json")))
    assert len(meta["frames"]) == 3 and np.array(meta["frames"][0]["transform_matrix"]).shape == (4, 4)
    rgba = imageio.imread(os.path.join(root, "train", "r_0.png"))
    assert rgba.shape == (16, 16, 4) and 0 < (rgba[..., 3] > 0).mean() < 1      # object and background both visible


def test_configargparse_standin(tmp_path):
    cap = _standin("configargparse")
    cfg = os.path.join(tmp_path, "c.txt")
    with open(cfg, "w") as fh:
        fh.write("expname = demo\nno_batching = True\nN_rand = 512  # comment\nlrate = 0.01\nhalf_res = False\n")
    p = cap.ArgumentParser()
    p.add_argument("--config", is_config_file=True)
    p.add_argument("--expname", type=str)
    p.add_argument("--no_batching", action="store_true")
    p.add_argument("--half_res", action="store_true")
    p.add_argument("--N_rand", type=int, default=4096)
    p.add_argument("--lrate", type=float, default=5e-4)
    a = p.parse_args(["--config", cfg, "--lrate", "0.02"])
    assert (a.expname, a.no_batching, a.half_res, a.N_rand, a.lrate) == ("demo", True, False, 512, 0.02)


def test_sample_pdf_bound_accepts_reorderings_and_rejects_wrong_bins():
    """oracle.sample_pdf_tolerance: zero error for the reference arithmetic, accepts a CDF summed in fp64 or by a
    Hillis-Steele scan (other summation orders), rejects samples displaced by two bins."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    rs = np.random.RandomState(3)
    R, S, Ni = 40, 64, 128
    z = np.sort(2.0 + 4.0 * rs.rand(R, S).astype(np.float32), -1)
    bins = torch.from_numpy(0.5 * (z[:, 1:] + z[:, :-1]))
    w = torch.from_numpy((rs.rand(R, S - 2).astype(np.float32)) ** 6)          # mostly ~0: pdf ~ 1e-5 / sum
    w[0] = 0.0
    w[1] = 0.0
    w[1, 30] = 1.0 - 62e-5                                                      # sum(w + 1e-5) == 1: denom at the switch
    u = torch.from_numpy(rs.rand(R, Ni).astype(np.float32))
    u[1, :30] = (torch.arange(30) + 0.5) * 1e-5                                 # inside the empty bins of row 1
    want, tol, chaotic = O.sample_pdf_tolerance(bins, w, u)
    assert torch.equal(want, O.sample_pdf(bins, w, u))
    assert bool(chaotic.any()), "the test is meant to contain branch-chaotic samples"

    def variant(cdf_fn):
        ww = w + 1e-5
        pdf = ww / ww.sum(-1, keepdim=True)
        cdf = torch.cat([torch.zeros(R, 1), cdf_fn(pdf)], -1)
        inds = torch.searchsorted(cdf, u.contiguous(), right=True)
        below, above = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
        c_lo, c_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
        b_lo, b_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
        denom = c_hi - c_lo
        denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
        return b_lo + (u - c_lo) / denom * (b_hi - b_lo)

    def scan(pdf):
        x = pdf.clone()
        d = 1
        while d < x.shape[1]:
            y = x.clone()
            y[:, d:] = x[:, d:] + x[:, :-d]
            x, d = y, d * 2
        return x

    for fn in (lambda p: torch.cumsum(p.double(), -1).float(), scan):
        got = variant(fn)
        assert bool(((got - want).abs() <= tol).all())
    width = (bins[:, 1:] - bins[:, :-1]).max()
    assert not bool((((want + 2.5 * width) - want).abs() <= tol).all())
