"""scene_prepare: wraps the reference-layout scene tensors into a ``surf_scene`` handle.

The reference hands raw tensors to every ``render()`` call (implicit_surface.py:268); the B200 path
converts them ONCE per scene into the compact HBM layout (int32 index tables, 1-bit masks, 32-byte
voxel rows, NHWC feature maps — DESIGN.md §3) and caches the handle keyed on the identity of the
input tensors, so the drop-in signatures stay unchanged.

Two parts with different lifetimes:
* the VOLUME part (sparse volumes, index tables, masks, matching volume; GBs, converted once per
  scene) — the cache key;
* the VIEW part (source images, feature pyramids, cameras; tens of MB, may change every step when
  ``SuRF.forward`` selects ``view_ids``, surf.py:140-146) — swapped in place with
  ``surf_scene_set_views`` when it changes; never forces a volume re-prepare.

Ownership: a ``PreparedScene`` frees its handle only when the last Python reference dies
(``__del__``).  The cache never destroys a scene a caller may still hold: eviction just drops the
cache's reference.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict
from typing import Optional, Sequence

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _ident(tensors):
    """Identity + version of a list of tensors (None allowed)."""
    return tuple(None if t is None else (id(t), t._version) for t in tensors)


def _as_list(x):
    return [] if x is None else ([x] if isinstance(x, torch.Tensor) else list(x))


class PreparedScene:
    """Owns a ``surf_scene*``.  Lists in renderer order (fine->coarse / high->low res)."""

    def __init__(self, volumes: Sequence[torch.Tensor], sparse_idxes: Sequence[torch.Tensor],
                 mask_volumes: Optional[Sequence[torch.Tensor]] = None,
                 matching_volume: Optional[torch.Tensor] = None, imgs: Optional[torch.Tensor] = None,
                 features: Optional[Sequence[torch.Tensor]] = None, intrs: Optional[torch.Tensor] = None,
                 c2ws: Optional[torch.Tensor] = None):
        lib = _lib.load()
        self._h = None
        self._lib = lib
        if isinstance(volumes, torch.Tensor):
            volumes, sparse_idxes = [volumes], [sparse_idxes]
            if mask_volumes is not None and isinstance(mask_volumes, torch.Tensor):
                mask_volumes = [mask_volumes]
        volumes, sparse_idxes = list(volumes), list(sparse_idxes)
        n_levels = len(volumes)
        if n_levels == 0 and matching_volume is None:
            raise ValueError("a scene needs volume levels or a matching volume")
        dev = volumes[0].device if n_levels else matching_volume.device
        if dev.type != "cuda":
            raise RuntimeError("surf_b200: scene tensors must live on a CUDA device (no CPU fallback)")
        self.device = dev
        if n_levels > _lib.MAX_LEVELS:
            raise ValueError("at most 4 volume levels supported")
        keep = []
        inp = _lib.SceneInputs()
        inp.n_levels = n_levels
        inp.feat_ch = int(volumes[0].shape[1]) if n_levels else 7      # 0 levels: matching-volume-only (MatchingField)
        for l in range(n_levels):
            v = _f32c(volumes[l])
            idx = sparse_idxes[l].detach()
            if idx.dtype != torch.int64:
                idx = idx.to(torch.int64)
            idx = idx.contiguous()
            if not (idx.dim() == 3 and idx.shape[0] == idx.shape[1] == idx.shape[2]):
                raise ValueError("sparse index tables must be cubic (N,N,N) (projector.py:324 assumes it)")
            keep += [v, idx]
            inp.dim[l] = int(idx.shape[0])
            inp.n_vox[l] = int(v.shape[0])
            inp.d_volumes[l] = v.data_ptr() if v.numel() else None
            inp.d_sparse_idx[l] = idx.data_ptr()
            if mask_volumes is not None:
                m = _f32c(mask_volumes[l])
                if m.numel() != idx.numel():
                    raise ValueError("mask volume %d does not match its index table" % l)
                keep.append(m)
                inp.d_mask_volumes[l] = m.data_ptr()
        if matching_volume is not None:
            mv = _f32c(matching_volume)
            keep.append(mv)
            inp.d_matching_volume = mv.data_ptr()
            inp.match_dim = int(mv.shape[-1])
            if not (mv.shape[-1] == mv.shape[-2] == mv.shape[-3]):
                raise ValueError("matching volume must be cubic")
        self.n_levels = n_levels
        self.n_views = 0
        self.n_src_views = 0
        self.has_matching = matching_volume is not None
        self.has_images = False
        self.intrs_host = self.c2ws_host = None
        self._view_key = None
        self._vol_versions = [v._version for v in volumes]
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.surf_scene_create(C.byref(inp), _stream(), C.byref(handle)), "scene_create")
            # conversion kernels read the inputs asynchronously: finish before `keep` may be freed
            torch.cuda.current_stream().synchronize()
        self._h = handle
        del keep
        if c2ws is not None:
            self.set_views(imgs, features, intrs, c2ws)
        elif imgs is not None:
            raise ValueError("imgs given without camera matrices")

    @classmethod
    def from_sparse(cls, coords_all, volumes_all, logits_all, dims_all):
        """Levels in BUILD order (coarse -> fine, dims doubling): coordinates (n,3), feature rows (n,feat_ch), matching
        logits (n,) / (n,1) or None.  Emits the prepared layout directly (surf_scene_create_sparse)."""
        lib = _lib.load()
        self = cls.__new__(cls)
        self._h = None
        self._lib = lib
        L = len(coords_all)
        dev = coords_all[0].device
        if dev.type != "cuda":
            raise RuntimeError("surf_b200: scene tensors must live on a CUDA device (no CPU fallback)")
        self.device = dev
        inp = _lib.SceneSparseInputs()
        inp.n_levels = L
        inp.feat_ch = int(volumes_all[0].shape[1])
        keep = []
        for b in range(L):
            co, vo = _f32c(coords_all[b]), _f32c(volumes_all[b])
            keep += [co, vo]
            inp.dim[b] = int(dims_all[b])
            inp.n_vox[b] = int(co.shape[0])
            inp.d_coords[b] = co.data_ptr()
            inp.d_volumes[b] = vo.data_ptr() if vo.numel() else None
            if logits_all is not None and logits_all[b] is not None:
                lg = _f32c(logits_all[b]).reshape(-1)
                keep.append(lg)
                inp.d_logits[b] = lg.data_ptr()
        self.n_levels = L
        self.n_views = 0
        self.n_src_views = 0
        self.has_matching = logits_all is not None and logits_all[-1] is not None
        self.has_images = False
        self.intrs_host = self.c2ws_host = None
        self._view_key = None
        self._vol_versions = [v._version for v in volumes_all]
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.surf_scene_create_sparse(C.byref(inp), _stream(), C.byref(handle)), "scene_create_sparse")
            torch.cuda.current_stream().synchronize()
        self._h = handle
        return self

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("scene already destroyed")
        return self._h

    def set_views(self, imgs, features, intrs, c2ws):
        """Installs / replaces the per-batch part (images, feature pyramids, cameras).  A no-op when the same
        tensor objects (same versions) are already installed."""
        key = _ident([imgs] + _as_list(features) + [intrs, c2ws])
        if key == self._view_key:
            return self
        if c2ws is None or intrs is None:
            raise ValueError("camera matrices (intrs, c2ws) are required")
        keep = []
        v = _lib.SceneViews()
        # host-side, with the reference's own routine (torch.inverse, projector.py:529)
        c2w_h = c2ws.detach().to(torch.float32).cpu().contiguous()
        w2c_h = torch.inverse(c2w_h).contiguous()
        K_h = intrs.detach().to(torch.float32).cpu().contiguous()
        keep += [c2w_h, w2c_h, K_h]
        v.h_c2ws, v.h_w2cs, v.h_intrs = c2w_h.data_ptr(), w2c_h.data_ptr(), K_h.data_ptr()
        v.n_views = int(c2w_h.shape[0])
        if imgs is not None:
            im = _f32c(imgs)
            keep.append(im)
            nv, _, H, W = im.shape
            if nv != v.n_views:
                raise ValueError("imgs has %d views, c2ws %d" % (nv, v.n_views))
            if features is None or len(features) != 4:
                raise ValueError("4 feature pyramid levels expected (high-res -> low-res)")
            v.img_h, v.img_w, v.n_feat_levels = H, W, 4
            v.d_imgs = im.data_ptr()
            for i, f in enumerate(features):
                f = _f32c(f)
                if tuple(f.shape) != (nv, 4, H >> i, W >> i):
                    raise ValueError("feature level %d has shape %s, expected %s"
                                     % (i, tuple(f.shape), (nv, 4, H >> i, W >> i)))
                keep.append(f)
                v.d_features[i] = f.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.surf_scene_set_views(self.handle, C.byref(v), _stream()), "scene_set_views")
            torch.cuda.current_stream().synchronize()      # `keep` (fp32 / contiguous copies) may be freed now
        self.n_views = v.n_views
        self.n_src_views = max(0, self.n_views - 1)
        self.has_images = self.has_images or imgs is not None
        self.intrs_host, self.c2ws_host = K_h, c2w_h         # the training extras form their 3x3 matrices from these
        self._view_key = key
        # the key holds ids: keep the objects alive so an id cannot be recycled by another tensor
        self._view_refs = [imgs] + _as_list(features) + [intrs, c2ws]
        return self

    def stats(self):
        st = _lib.SceneStats()
        _lib.check(self._lib.surf_scene_get_stats(self.handle, C.byref(st)), "scene_get_stats")
        return {"bytes_index": st.bytes_index, "bytes_volumes": st.bytes_volumes, "bytes_masks": st.bytes_masks,
                "bytes_matching": st.bytes_matching, "bytes_images": st.bytes_images,
                "n_vox": [int(st.n_vox[i]) for i in range(self.n_levels)]}

    def update_volume(self, level: int, volume: torch.Tensor):
        """Refreshes the padded copy of ``volumes[level]`` (finetune optimises the volumes in place, surf.py:43-44)."""
        v = _f32c(volume)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.surf_scene_update_volume(self.handle, level, v.data_ptr(), v.shape[0], _stream()),
                       "scene_update_volume")
            torch.cuda.current_stream().synchronize()

    def destroy(self):
        """Frees the device memory now.  Only for owners that know nobody else holds the scene."""
        if getattr(self, "_h", None) is not None:
            self._lib.surf_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class SceneCache:
    """Small LRU of PreparedScene keyed on the identity of the VOLUME-part tensors.

    An entry is only valid while every keyed tensor object is still alive (weak references): a freed
    tensor's address / id can be reused, so ids alone would alias different scenes.  In-place updates of
    ``volumes[l]`` (finetune) refresh just that level; any other in-place change re-prepares.  Eviction
    and invalidation only drop the cache's reference — a scene object a caller still holds stays valid."""

    def __init__(self, capacity=4):
        self.capacity = capacity
        self._d = OrderedDict()

    def get(self, volumes, sparse_idxes, mask_volumes=None, matching_volume=None, imgs=None, features=None,
            intrs=None, c2ws=None) -> PreparedScene:
        vols, idxs = _as_list(volumes), _as_list(sparse_idxes)
        rest = _as_list(mask_volumes) + [matching_volume]
        flat = vols + idxs + rest
        key = tuple(None if t is None else id(t) for t in flat)
        hit = self._d.get(key)
        sc = None
        if hit is not None:
            refs, versions, cand = hit
            alive = all((r is None and t is None) or (r is not None and r() is t) for r, t in zip(refs, flat))
            fixed_ok = alive and versions[len(vols):] == tuple(None if t is None else t._version for t in flat[len(vols):])
            if fixed_ok and cand._h is not None:
                sc = cand
                for l, v in enumerate(vols):          # volumes optimised in place: refresh the padded copies
                    if v._version != versions[l]:
                        sc.update_volume(l, v)
                self._d[key] = (refs, tuple(None if t is None else t._version for t in flat), sc)
                self._d.move_to_end(key)
            else:
                del self._d[key]
        if sc is None:
            sc = PreparedScene(volumes, sparse_idxes, mask_volumes, matching_volume)
            refs = tuple(None if t is None else weakref.ref(t) for t in flat)
            versions = tuple(None if t is None else t._version for t in flat)
            self._d[key] = (refs, versions, sc)
            while len(self._d) > self.capacity:
                self._d.popitem(last=False)           # drop our reference only; __del__ frees when unused
        if c2ws is not None:
            sc.set_views(imgs, features, intrs, c2ws)
        elif imgs is not None:
            raise ValueError("imgs given without camera matrices")
        return sc

    def clear(self):
        self._d.clear()


GLOBAL_SCENE_CACHE = SceneCache()
# throw-away helper scenes (mask-only / image-only stage calls of modules/projector.py) never share the LRU of
# the render scenes
AUX_SCENE_CACHE = SceneCache(capacity=2)
