"""Mirrors of the live functions of models/modules/projector.py, bound to the stage-isolated C-ABI
entry points (the same device code the fused render path runs)."""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from .. import _lib
from ..scene import AUX_SCENE_CACHE, GLOBAL_SCENE_CACHE, PreparedScene


_MASK_SCENES = {}     # at most one mask-only helper scene


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pts(pts):
    return pts.detach().reshape(-1, 3).to(torch.float32).contiguous()


def lookup_volume(pts, volume, sample_mode="nearest"):
    """projector.py:392-420 for the mask volumes: returns (n,1) float = any-level nearest lookup.

    Only the mask use (``sample_mode='nearest'`` followed by ``.any(-1)``, implicit_surface.py:86)
    has a stage entry point; the bilinear probe of the matching volume lives inside
    ``surf_sample_rays``.  ``volume`` = PreparedScene or the list of (1,1,N,N,N) mask volumes."""
    if sample_mode != "nearest":
        raise NotImplementedError("only the nearest (mask) lookup is exposed as a stage")
    scene = volume if isinstance(volume, PreparedScene) else _mask_only_scene(volume)
    p = _pts(pts)
    out = torch.empty(p.shape[0], dtype=torch.uint8, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().surf_point_mask(scene.handle, p.data_ptr(), p.shape[0], out.data_ptr(), _stream()),
                   "point_mask")
    return out.to(torch.float32)[:, None]


def _mask_only_scene(mask_volumes):
    if isinstance(mask_volumes, torch.Tensor):
        mask_volumes = [mask_volumes]
    dev = mask_volumes[0].device
    vols = [torch.zeros((0, 7), device=dev) for _ in mask_volumes]
    idxs = [torch.full(tuple(m.shape[-3:]), -1, dtype=torch.int64, device=dev) for m in mask_volumes]
        # a throw-away scene: keyed on the mask tensors only, kept out of the render-scene LRU
    key = tuple(id(m) for m in mask_volumes)
    hit = _MASK_SCENES.get(key)
    if hit is not None and all(r() is m for r, m in zip(hit[0], mask_volumes)) and \
            hit[1] == tuple(m._version for m in mask_volumes):
        return hit[2]
    sc = PreparedScene(vols, idxs, list(mask_volumes))
    _MASK_SCENES.clear()
    _MASK_SCENES[key] = (tuple(weakref.ref(m) for m in mask_volumes), tuple(m._version for m in mask_volumes), sc)
    return sc


_IMG_SCENES = {}      # one volume-less helper scene per device; its views are swapped per call


def _image_only_scene(dev):
    sc = _IMG_SCENES.get(dev)
    if sc is None:
        vols = [torch.zeros((0, 7), device=dev)]
        idxs = [torch.full((2, 2, 2), -1, dtype=torch.int64, device=dev)]
        sc = _IMG_SCENES[dev] = PreparedScene(vols, idxs)
    return sc


def lookup_sparse_volume(pts, sparse_volumes, sparse_idxs=None):
    """projector.py:377-390 -> (n, 7 * levels), levels in the order given (fine -> coarse)."""
    scene = sparse_volumes if isinstance(sparse_volumes, PreparedScene) else AUX_SCENE_CACHE.get(
        sparse_volumes, sparse_idxs)
    p = _pts(pts)
    out = torch.empty((p.shape[0], 7 * scene.n_levels), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().surf_lookup_sparse(scene.handle, p.data_ptr(), p.shape[0], out.data_ptr(), _stream()),
                   "lookup_sparse")
    return out


def lookup_feature(pts, imgs, intrs=None, c2ws=None, features=None):
    """projector.py:501-556 -> feat_views (n,V,19), ray_diff (n,V,4), mask (n,V) bool.
    ``imgs`` may be a PreparedScene that was built with images."""
    if isinstance(imgs, PreparedScene):
        scene = imgs
    else:
        dev = imgs.device
        scene = _image_only_scene(dev).set_views(imgs, features, intrs, c2ws)
    p = _pts(pts)
    n, V = p.shape[0], scene.n_src_views
    fv = torch.empty((n, V, 19), dtype=torch.float32, device=p.device)
    rd = torch.empty((n, V, 4), dtype=torch.float32, device=p.device)
    m = torch.empty((n, V), dtype=torch.uint8, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().surf_lookup_feature(scene.handle, p.data_ptr(), n, fv.data_ptr(), rd.data_ptr(),
                                                   m.data_ptr(), _stream()), "lookup_feature")
    return fv, rd, m.bool()
