"""surf_b200 — B200-native (sm_100a) implementation of SuRF's per-ray volume-rendering hot path.

Python host code (this package) mirrors the reference's module API; all compute runs in hand-written
CUDA kernels (csrc/*.cu) behind the C ABI declared in include/surf_b200.h.  No CPU fallback.
"""
from .conf import ConfigTree, default_implicit_surface_conf, parse_file, parse_string  # noqa: F401

__version__ = "0.1.0"


def build(force=False, verbose=False):
    from .build import build as _b
    return _b(force=force, verbose=verbose)
