"""Executes the reference's UNMODIFIED run_nerf.py (the git-ignored copy under oracle/_ref made by oracle/make_ref.py,
or /root/reference) against this repository's drop-in modules: the two-line launcher of INTEGRATION.md plus the
offline stand-ins of SURVEY App. B11.

    python tests/harness/run_reference_main.py <workdir> [--iters 50] [extra run_nerf.py options]

sys.path order: stand-ins (configargparse, imageio, matplotlib, cv2, tqdm -- third-party packages that are not
installed here) < hashnerf-pytorch_b200/ (run_nerf_helpers, ray_util, loss, radam, models, embedding: OURS) <
the reference tree (run_nerf.py, util.py, bbox.py, load/*: THEIRS, unmodified).  run_nerf.py is run as __main__,
so its own ``torch.set_default_tensor_type('torch.cuda.FloatTensor')`` (:725) and argument parsing execute.
TEST HARNESS ONLY."""
from __future__ import annotations

import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def reference_root() -> str:
    for cand in (os.environ.get("HASHNERF_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "run_nerf.py")):
            return cand
    raise SystemExit("reference run_nerf.py not found (run oracle/make_ref.py where /root/reference exists)")


def main(argv):
    workdir = os.path.abspath(argv[0])
    rest = list(argv[1:])
    iters = 50
    if "--iters" in rest:
        k = rest.index("--iters")
        iters = int(rest[k + 1])
        del rest[k:k + 2]
    seed = None
    if "--seed" in rest:       # the script seeds numpy only (run_nerf.py:733): fix the parameter initialisation too
        k = rest.index("--seed")
        seed = int(rest[k + 1])
        del rest[k:k + 2]
    os.makedirs(workdir, exist_ok=True)
    sys.path.insert(0, os.path.join(HERE))
    import make_scene
    scene = make_scene.make(os.path.join(workdir, "scene"))
    cfg = os.path.join(workdir, "harness.txt")
    with open(cfg, "w") as fh:   # chair.txt's options at harness size
        fh.write("\n".join([
            "expname = harness", f"basedir = {os.path.join(workdir, 'logs')}", f"datadir = {scene}",
            "dataset_type = blender", "no_batching = True", "use_viewdirs = True", "white_bkgd = True",
            "lrate_decay = 500", "N_samples = 32", "N_importance = 64", "N_rand = 512", "precrop_iters = 10",
            "precrop_frac = 0.5", "testskip = 1", "log2_hashmap_size = 14", "finest_res = 256", "lrate = 0.01",
            f"i_print = 10", f"i_weights = {iters}", f"i_testset = {iters}", f"i_video = {iters}", "chunk = 8192"]
            + (["perturb = 0."] if os.environ.get("HN_HARNESS_NO_PERTURB") == "1" else []) + [""]))
    ref = reference_root()
    os.environ["HN_HARNESS_ITERS"] = str(iters)
    sys.path[:0] = [os.path.join(HERE, "standins"), os.path.join(ROOT, "hashnerf-pytorch_b200"), ref]
    sys.argv = [os.path.join(ref, "run_nerf.py"), "--config", cfg] + rest
    os.chdir(workdir)
    if seed is not None:
        import torch
        torch.manual_seed(seed)
    try:
        runpy.run_path(os.path.join(ref, "run_nerf.py"), run_name="__main__")
    finally:
        if os.environ.get("HN_AUTO_GRAPH") == "1":   # how often render_rays was replayed as CUDA graphs under the script
            from hn_b200 import autograph
            print("[autograph]", dict(autograph.stats), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
