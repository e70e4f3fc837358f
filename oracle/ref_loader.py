"""Load the UNMODIFIED reference hot-path sources from /root/reference for golden-vector generation.

TEST INFRASTRUCTURE ONLY.  Nothing here is product code; nothing here may be imported by the
package under ``hashnerf-pytorch_b200/``.  It is used by ``oracle/gen_golden.py`` (run in the
build container, where ``/root/reference`` exists) to produce the fixtures under
``tests/golden/`` and by CPU tests that pin ``oracle/oracle.py`` against the live reference
when the reference tree happens to be present.  The GPU box has no ``/root/reference``: every
caller must check :func:`available` first.

No reference source is copied into this repository's history: the files are read where they lie
(``/root/reference``, or the git-ignored copy ``oracle/_ref/`` that ``oracle/make_ref.py`` makes so that the same
loader works on the GPU box) and exec'd with the import-time patches listed in SURVEY.md section 8c:

* ``embedding/hash_encoding.py:10-11`` allocates ``BOX_OFFSETS`` with ``device='cuda'``;
  the string is replaced by the requested device so the module imports without a driver.
* ``run_nerf_helpers.py:1,4,8,15-21`` import matplotlib/pdb/tqdm/itself/radam/ray_util and two
  names that do not exist where they are looked for (Appendix B1/B2); those lines are blanked
  and the modules they wanted are injected into the namespace instead.
* ``loss.py:8`` imports ``hash`` from the un-importable module; blanked and injected.
"""
from __future__ import annotations

import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # made by oracle/make_ref.py


def _find_root() -> str:
    """The live checkout if there is one (build container), else the git-ignored copy that travels to the GPU box."""
    env = os.environ.get("HASHNERF_REFERENCE_ROOT")
    for cand in (env, "/root/reference", _VENDORED):
        if cand and os.path.isfile(os.path.join(cand, "embedding", "hash_encoding.py")):
            return cand
    return env or "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "embedding", "hash_encoding.py"))


def _read(rel: str) -> list[str]:
    with open(os.path.join(REF_ROOT, rel), "r") as fh:
        return fh.read().split("\n")


def _exec_module(name: str, rel: str, lines: list[str], inject: dict | None = None) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF_ROOT, rel)
    if inject:
        mod.__dict__.update(inject)
    code = compile("\n".join(lines), mod.__file__, "exec")
    sys.modules[name] = mod  # @dataclass resolves cls.__module__ through sys.modules
    exec(code, mod.__dict__)
    return mod


_CACHE: dict[str, types.SimpleNamespace] = {}


def load(device: str = "cpu") -> types.SimpleNamespace:
    """Return a namespace with the reference's hot-path callables bound to ``device``."""
    if device in _CACHE:
        return _CACHE[device]
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")

    # embedding/hash_encoding.py -- patch the import-time CUDA allocation only.
    src = _read("embedding/hash_encoding.py")
    src = [ln.replace("device='cuda'", f"device='{device}'") for ln in src]
    hash_mod = _exec_module("_ref_hash_encoding", "embedding/hash_encoding.py", src)

    # These two import cleanly as files; load them by path so sys.path is left alone.
    sh_mod = _exec_module("_ref_spherical_harmonic", "embedding/spherical_harmonic.py",
                          _read("embedding/spherical_harmonic.py"))
    models_src = _read("models.py")
    models_mod = _exec_module("_ref_models", "models.py", models_src)
    radam_mod = _exec_module("_ref_radam", "radam.py", _read("radam.py"))

    helpers_src = _read("run_nerf_helpers.py")
    blank = {1, 4, 8, 15, 16, 17, 18, 19, 20, 21}  # 1-based line numbers
    helpers_src = [("" if (i + 1) in blank else ln) for i, ln in enumerate(helpers_src)]
    helpers_mod = _exec_module(
        "_ref_run_nerf_helpers", "run_nerf_helpers.py", helpers_src,
        inject=dict(HashEmbedder=hash_mod.HashEmbedder, SHEncoder=sh_mod.SHEncoder,
                    NeRFSmall=models_mod.NeRFSmall, NeRF=models_mod.NeRF,
                    RAdam=radam_mod.RAdam))

    # ray_util.py imports kornia (not installed) for the st3d helpers only; blank that line (:3)
    ray_src = _read("ray_util.py")
    ray_src = [("" if (i + 1) == 3 else ln) for i, ln in enumerate(ray_src)]
    ray_mod = _exec_module("_ref_ray_util", "ray_util.py", ray_src)

    loss_src = _read("loss.py")
    loss_src = [("" if (i + 1) == 8 else ln) for i, ln in enumerate(loss_src)]
    loss_mod = _exec_module("_ref_loss", "loss.py", loss_src, inject=dict(hash=hash_mod.hash))

    ns = types.SimpleNamespace(
        device=device,
        hash_encoding=hash_mod, spherical_harmonic=sh_mod, models=models_mod,
        helpers=helpers_mod, loss=loss_mod, radam=radam_mod,
        HashEmbedder=hash_mod.HashEmbedder, hash=hash_mod.hash,
        trilinear_interp=hash_mod.trilinear_interp,
        SHEncoder=sh_mod.SHEncoder, NeRFSmall=models_mod.NeRFSmall,
        run_network=helpers_mod.run_network, render_rays=helpers_mod.render_rays,
        raw2outputs=helpers_mod.raw2outputs, sample_pdf=helpers_mod.sample_pdf,
        total_variation_loss=loss_mod.total_variation_loss, RAdam=radam_mod.RAdam,
        ray_util=ray_mod, get_rays=ray_mod.get_rays, get_rays_np=ray_mod.get_rays_np,
        get_ndc_rays=ray_mod.get_ndc_rays,
    )
    _CACHE[device] = ns
    return ns


if __name__ == "__main__":  # quick self-check in the build container
    import torch

    ref = load("cpu")
    bbox = (torch.tensor([-1.5, -1.5, -1.5]), torch.tensor([1.5, 1.5, 1.5]))
    emb = ref.HashEmbedder(bbox, log2_hashmap_size=14)
    y, keep = emb(torch.rand(5, 3) * 3 - 1.5)
    print("reference HashEmbedder ok:", tuple(y.shape), keep.tolist())
