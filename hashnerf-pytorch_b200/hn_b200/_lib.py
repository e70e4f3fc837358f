"""ctypes binding of libhashnerf_b200.so (C ABI declared in include/hashnerf_b200.h).

There is deliberately no fallback: if the shared object is missing or a call fails, a
RuntimeError is raised.  Nothing in this package can execute the hot path on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libhashnerf_b200.so")

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_d = C.c_double

ABI_VERSION = 2  # HN_ABI_VERSION of include/hashnerf_b200.h

# name -> (restype, argtypes); mirrors include/hashnerf_b200.h one to one
SIGNATURES = {
    "hn_abi_version": (_i, []),
    "hn_last_error_string": (C.c_char_p, []),
    "hn_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "hn_set_tuning": (_i, [C.c_char_p, _i]),
    "hn_spatial_hash": (_i, [_p, _l, _i, _i, _p, _p]),
    "hn_voxel_vertices": (_i, [_p, _p, _p, _l, _i, _i, _p, _p, _p, _p]),
    "hn_hash_encode_fwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _p, _p]),
    "hn_hash_encode_bwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "hn_hash_encode_bwd_ordered": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "hn_hash_sort_workspace_bytes": (_l, [_l, _i]),
    "hn_hash_sort_points": (_i, [_p, _p, _l, _i, _p, _p, _p]),
    "hn_hash_encode_fwd_sorted": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _p, _p]),
    "hn_hash_encode_bwd_sorted": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "hn_hash_encode_bwd_sorted_levels": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _i, _i, _p]),
    "hn_sh_encode": (_i, [_p, _l, _i, _p, _p]),
    "hn_mlp_fwd": (_i, [_p, _l, _p, _l, _l, _p, _p, _l, _p, _p, _p]),
    "hn_mlp_bwd_workspace_bytes": (_l, [_l]),
    "hn_mlp_bwd": (_i, [_p, _l, _p, _l, _l, _p, _p, _p, _p, _l, _p, _p, _p, _p]),
    "hn_composite_fwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "hn_composite_bwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hn_sample_pdf": (_i, [_p, _p, _p, _p, _l, _i, _i, _p, _p]),
    "hn_resample": (_i, [_p, _p, _p, _p, _l, _i, _i, _p, _p, _p, _p]),
    "hn_sort_concat_rows": (_i, [_p, _i, _p, _i, _l, _p, _p]),
    "hn_coarse_z": (_i, [_p, _p, _l, _p, _p, _l, _i, _i, _p, _p]),
    "hn_ray_points": (_i, [_p, _p, _l, _p, _l, _i, _p, _p]),
    "hn_get_rays": (_i, [_i, _i, _f, _f, _f, _f, _p, _l, _p, _p]),
    "hn_pack_rays": (_i, [_p, _l, _p, _l, _p, _l, _f, _f, _l, _p, _p]),
    "hn_tv_loss_fwd": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "hn_tv_loss_bwd": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "hn_tv_loss_fwd_levels": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "hn_tv_loss_bwd_levels": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "hn_ndc_rays": (_i, [_i, _i, _d, _d, _p, _l, _p, _l, _l, _p, _p, _p]),
    "hn_sample_rays": (_i, [_p, _i, _p, _i, _i, _f, _f, _f, _f, _f, _f, _p, _l, _p, _p, _p, _p]),
    "hn_mse_fwd": (_i, [_p, _p, _l, _p, _p]),
    "hn_mse_bwd": (_i, [_p, _p, _l, _p, _p, _p, _p]),
    "hn_dp_barrier": (_i, [_p, _i, _i, _i, C.c_uint32, _p]),
    "hn_dp_reduce_update": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _l, _p, _p]),
    "hn_radam_step_dev": (_i, [_p, _p, _p, _p, _l, _p, _p]),
    "hn_radam_step": (_i, [_p, _p, _p, _p, _l, _f, _f, _f, _d, _d, _d, _i, _f, _i, _p]),
}

# kernels launched per entry point (1 unless listed): hn_hash_sort_points = histogram + scan + partition + local sort
# (two-level form; its memset node is not counted)
KERNELS_PER_CALL = {"hn_hash_sort_points": 4, "hn_mlp_bwd": 3}  # hn_mlp_bwd: image prep + fused kernel + dW row sum

_lib = None
launches = 0  # number of CUDA kernels launched through call() (bench.py reports it as gpu_launches)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python hashnerf-pytorch_b200/hn_b200/build.py` "
            "(or __graft_entry__.build()).  This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.hn_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libhashnerf_b200 ABI {lib.hn_abi_version()} != {ABI_VERSION} expected by the Python shims")
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise RuntimeError on a non-zero status."""
    global launches
    lib = _lib if _lib is not None else load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.hn_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed with status {rc}: {msg}")
    launches += KERNELS_PER_CALL.get(name, 1)


def set_tuning(key: str, value: int) -> None:
    lib = load()
    rc = lib.hn_set_tuning(key.encode(), int(value))
    if rc != 0:
        raise RuntimeError(lib.hn_last_error_string().decode())
