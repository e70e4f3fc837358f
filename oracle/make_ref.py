"""Recipe that vendors the reference checkout into ``oracle/_ref/`` (git-ignored; it travels to the GPU box
with the tree, like the built ``.so``).

TEST / BASELINE INFRASTRUCTURE ONLY.  No reference source enters this repository's history: the files are copied,
unmodified, from ``/root/reference`` (or ``$HASHNERF_REFERENCE_ROOT``) into a directory that ``.gitignore`` lists.
``oracle/ref_loader.py`` reads the copy when ``/root/reference`` itself is absent (the GPU box), which is what lets
``bench.py --impl reference`` / ``cpu_baseline`` time the reference's OWN PyTorch code (``kind: "reference"``),
lets the ``reference_gpu`` leg run that same code on ``device='cuda'`` beside ours, and lets
``tests/harness/run_reference_main.py`` execute the reference's unmodified ``run_nerf.py``.

    python oracle/make_ref.py            # called by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
# everything run_nerf.py imports from its own tree + the two configs the bench shapes come from
FILES = [
    "embedding/__init__.py", "embedding/hash_encoding.py", "embedding/spherical_harmonic.py", "embedding/embedder.py",
    "models.py", "run_nerf_helpers.py", "run_nerf.py", "loss.py", "radam.py", "optimizer.py", "ray_util.py", "util.py",
    "bbox.py", "load/load_blender.py", "load/load_llff.py", "load/load_deepvoxels.py", "load/load_LINEMOD.py",
    "load/load_scannet.py", "load/load_st3d.py", "configs/chair.txt", "configs/fern.txt", "LICENSE",
]


def source_root() -> str:
    return os.environ.get("HASHNERF_REFERENCE_ROOT", "/root/reference")


def make(verbose: bool = False) -> str | None:
    src = source_root()
    if not os.path.isfile(os.path.join(src, "embedding", "hash_encoding.py")):
        return None
    os.makedirs(DEST, exist_ok=True)
    for rel in FILES:
        s = os.path.join(src, rel)
        if not os.path.isfile(s):
            continue  # optional files (e.g. embedding/__init__.py does not exist upstream)
        d = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        if verbose:
            print("copied", rel)
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as fh:
        fh.write("Unmodified copies of files of mache102/HashNeRF-pytorch made by oracle/make_ref.py.\n"
                 "Not part of this repository (git-ignored); used as checker / CPU and same-GPU baseline only.\n")
    return DEST


if __name__ == "__main__":
    out = make(verbose=True)
    print(out if out else f"reference tree not found under {source_root()}", file=sys.stderr if not out else sys.stdout)
    sys.exit(0 if out else 1)
