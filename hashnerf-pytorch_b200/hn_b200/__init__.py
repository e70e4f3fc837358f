"""hn_b200 -- host-side glue between the reference's Python API and libhashnerf_b200.so.

``ops`` holds thin wrappers (pointer + shape marshalling, autograd Functions); the modules one
directory up (``embedding/hash_encoding.py``, ``models.py``, ``run_nerf_helpers.py`` ...) mirror the
reference's module names so an unmodified ``run_nerf.py`` imports them.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
