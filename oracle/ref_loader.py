"""Loader for the UNMODIFIED reference (prstrive/SuRF) hot-path modules.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (in the build
container, where ``/root/reference`` exists) to generate the golden vectors
under ``tests/golden/`` that pin ``oracle/surf_oracle.py``.  Nothing here is
imported by the product (``surf_b200/``), by the ``-m gpu`` tests, by
``smoke()`` or by ``bench.py``: ``/root/reference`` does not exist on the GPU
box.

Three shims are needed to import ``models/modules/implicit_surface.py`` on a
CPU-only box without its un-vendored dependencies:
  * ``mcubes`` (PyMCubes 0.1.4, not installed) is imported at module top
    (implicit_surface.py:5) -> empty stub module;
  * ``models.modules.grid_sample_cuda.cuda_gridsample`` JIT-compiles a CUDA
    extension at import (projector.py:5, cuda_gridsample.py:5) -> stub exposing
    ``grid_sample_3d`` (only reachable via the dead sample_mode="grad" path);
  * a hard ``.cuda()`` at implicit_surface.py:189 -> identity on CPU.
No reference source is copied; the modules are imported from where they lie.
"""
import os
import sys
import types

import torch


def find_reference():
    for cand in (os.environ.get("SURF_REF"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"),
                 "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "models", "modules")):
            return cand
    return None


def load_reference():
    """Returns the reference ``models.modules.implicit_surface`` module."""
    ref = find_reference()
    if ref is None:
        raise RuntimeError("reference tree not found (set SURF_REF)")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))
    name = "models.modules.grid_sample_cuda"
    if name + ".cuda_gridsample" not in sys.modules:
        cg = types.ModuleType(name + ".cuda_gridsample")

        def grid_sample_3d(inp, grid, padding_mode="zeros", align_corners=True):
            return torch.nn.functional.grid_sample(
                inp, grid, mode="bilinear", padding_mode=padding_mode, align_corners=align_corners)

        cg.grid_sample_3d = grid_sample_3d
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        pkg.cuda_gridsample = cg
        sys.modules[name] = pkg
        sys.modules[name + ".cuda_gridsample"] = cg
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    from models.modules import implicit_surface  # noqa: E402  (reference module)
    return implicit_surface


class DictConf(dict):
    """Minimal pyhocon-like view over nested dicts (get_list/get_int/get_float + [])."""

    def _walk(self, key):
        cur = self
        for part in key.split("."):
            cur = cur[part]
        return cur

    def get_list(self, key, default=None):
        try:
            return list(self._walk(key))
        except KeyError:
            return default

    def get_int(self, key, default=None):
        try:
            return int(self._walk(key))
        except KeyError:
            return default

    def get_float(self, key, default=None):
        try:
            return float(self._walk(key))
        except KeyError:
            return default

    def get_bool(self, key, default=None):
        try:
            return bool(self._walk(key))
        except KeyError:
            return default
