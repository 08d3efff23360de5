"""Volume — drop-in for models/modules/volume.py:7-168: the producers of the scene tensors.

Same constructor conf (``base_volume_dim``, ``bounding``), same stateful ``volume_dim`` / ``voxel_size`` progression
(``init_coords`` resets, ``up_sample`` doubles), same ``agg_mlp`` parameter names.  The compute (projection of every
voxel into every view with multi-scale bilinear gathers + the view-aggregation MLP, the depth-consistency filter, the
trilinear 2x up-sampling) runs in csrc/volume.cu; ``init_coords`` / ``up_sample`` / ``sparse2dense`` / ``get_index``
return the REFERENCE layouts (they are pure data movement, done with torch indexing on the device) for callers that
want them — the render path does not: ``to_prepared_scene`` hands coordinates / features / logits of all levels to
``surf_scene_create_sparse``, which emits the compact layout (int32 index, 1-bit masks, fp32 matching volume) directly.
Outputs are detached (no autograd through the kernels).  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ..scene import PreparedScene


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class Volume(nn.Module):
    def __init__(self, confs):
        super().__init__()
        self.base_volume_dim = confs.get_list("base_volume_dim")
        self.bounding = np.array(confs.get_list("bounding", default=[[-1, 1], [-1, 1], [-1, 1]]))
        self.origin = self.bounding[:, 0]
        self.agg_mlp = nn.Sequential(nn.Linear(4, 8), nn.ELU(inplace=True), nn.Linear(8, 1))

    # -- data movement in the reference layouts ------------------------------------------------------
    def init_coords(self):
        """volume.py:21-33."""
        self.volume_dim = np.array(self.base_volume_dim)
        self.voxel_size = (self.bounding[:, 1] - self.bounding[:, 0]) / (self.volume_dim - 1)
        g = torch.stack(torch.meshgrid(*[torch.arange(0, int(d)) for d in self.volume_dim], indexing="ij")).float()
        return g.view(3, -1).permute(1, 0).contiguous()

    def up_sample(self, pre_coords, pre_feat, num=8):
        """volume.py:35-52 (doubles ``pre_coords`` in place like the reference)."""
        self.volume_dim *= 2
        self.voxel_size = (self.bounding[:, 1] - self.bounding[:, 0]) / (self.volume_dim - 1)
        with torch.no_grad():
            pre_coords *= 2
            offs = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]],
                                dtype=pre_coords.dtype, device=pre_coords.device)[:num]
            up_coords = (pre_coords[:, None, :] + offs[None]).reshape(-1, 3).contiguous()
            up_feat = pre_feat[:, None, :].expand(-1, num, -1).reshape(-1, pre_feat.shape[1]).contiguous()
        return up_coords, up_feat

    def sparse2dense(self, feats, coords, pre_volume):
        """volume.py:99-121, reference layout: (1,c,D,H,W) volume + (1,1,D,H,W) fp32 mask."""
        dims = [int(d) for d in self.volume_dim]
        c = feats.shape[1]
        dev = feats.device
        dense = torch.zeros([1, dims[0], dims[1], dims[2], c], device=dev)
        if pre_volume is not None:
            dense[..., :1] = self.upsample2x(pre_volume).permute(0, 2, 3, 4, 1)
        maskv = torch.zeros([1, dims[0], dims[1], dims[2], 1], device=dev)
        loc = coords.to(torch.int64)
        dense[:, loc[:, 0], loc[:, 1], loc[:, 2]] = feats.detach()
        maskv[:, loc[:, 0], loc[:, 1], loc[:, 2]] = 1.0
        return dense.permute(0, 4, 1, 2, 3), maskv.permute(0, 4, 1, 2, 3)

    def get_index(self, coords):
        """volume.py:123-132, reference layout: int64 table, -1 = empty."""
        dims = [int(d) for d in self.volume_dim]
        t = torch.full(dims, -1, dtype=torch.int64, device=coords.device)
        loc = coords.to(torch.int64)
        t[loc[:, 0], loc[:, 1], loc[:, 2]] = torch.arange(coords.shape[0], dtype=torch.int64, device=coords.device)
        return t

    # -- kernels -----------------------------------------------------------------------------------------
    @staticmethod
    def upsample2x(volume):
        """F.interpolate(volume (1,1,D,H,W), scale_factor=2, mode='trilinear') (volume.py:108)."""
        v = _f32c(volume)
        _, _, D, H, W = v.shape
        out = torch.empty((1, 1, 2 * D, 2 * H, 2 * W), dtype=torch.float32, device=v.device)
        with torch.cuda.device(v.device):
            _lib.check(_lib.load().surf_volume_upsample2x(v.data_ptr(), D, H, W, out.data_ptr(), _stream()), "upsample2x")
        return out

    def _views(self, intrs, c2ws, norm_h, norm_w, keep):
        c2w = c2ws.detach().float().cpu().contiguous()
        w2c = torch.inverse(c2w).contiguous()
        K = intrs.detach().float().cpu().contiguous()
        keep += [w2c, K]
        vw = _lib.VolumeViews()
        vw.n_views, vw.norm_h, vw.norm_w = int(K.shape[0]), int(norm_h), int(norm_w)
        vw.h_w2cs, vw.h_intrs = w2c.data_ptr(), K.data_ptr()
        return vw

    def _vs_org(self, keep):
        # float32 of the float64 numpy values, like torch.tensor(self.voxel_size).type_as(coords) (volume.py:65)
        vs = torch.tensor(self.voxel_size).to(torch.float32).contiguous()
        org = torch.tensor(self.origin).to(torch.float32).contiguous()
        keep += [vs, org]
        return vs, org

    def back_proj_multiscale(self, feats, coords, intrs, c2ws, stage_idx):
        """volume.py:54-97 -> (feat_vol (n,8), mask_vol (n,) bool)."""
        lib = _lib.load()
        keep = []
        co = _f32c(coords)
        dev = co.device
        n = co.shape[0]
        fl = [_f32c(f) for f in feats[stage_idx:]]
        nv, c, h, w = feats[-1].shape
        vw = self._views(intrs, c2ws, h, w, keep)
        vs, org = self._vs_org(keep)
        ptrs = (C.c_void_p * len(fl))(*[f.data_ptr() for f in fl])
        hs = (C.c_int32 * len(fl))(*[int(f.shape[2]) for f in fl])
        ws = (C.c_int32 * len(fl))(*[int(f.shape[3]) for f in fl])
        sd = self.agg_mlp.state_dict()
        agg = torch.cat([sd["0.weight"].reshape(-1), sd["0.bias"].reshape(-1), sd["2.weight"].reshape(-1),
                         sd["2.bias"].reshape(-1)]).detach().float().cpu().contiguous()
        fv = torch.empty((n, 2 * c), dtype=torch.float32, device=dev)
        mk = torch.empty((n,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.surf_volume_back_proj(C.byref(vw), ptrs, hs, ws, len(fl), int(c), agg.data_ptr(), co.data_ptr(), n,
                                                 vs.data_ptr(), org.data_ptr(), fv.data_ptr(), mk.data_ptr(), _stream()),
                       "volume_back_proj")
            torch.cuda.current_stream().synchronize()          # `fl` copies may be freed
        return fv, mk.bool()

    def depth_filtering(self, depths, coords, feats, intrs, c2ws, depth_range):
        """volume.py:134-168 -> (valid_coords, valid_feats)."""
        lib = _lib.load()
        keep = []
        d = _f32c(torch.stack(list(depths), dim=0))
        nv, h, w = d.shape
        co = _f32c(coords)
        vw = self._views(intrs, c2ws, h, w, keep)
        vs, org = self._vs_org(keep)
        valid = torch.empty((co.shape[0],), dtype=torch.uint8, device=co.device)
        with torch.cuda.device(co.device):
            _lib.check(lib.surf_volume_depth_filter(C.byref(vw), d.data_ptr(), co.data_ptr(), co.shape[0], vs.data_ptr(),
                                                    org.data_ptr(), float(depth_range), valid.data_ptr(), _stream()),
                       "volume_depth_filter")
        m = valid.bool()
        return coords[m], feats[m]

    # -- the compact hand-off to the render path ------------------------------------------------------
    @staticmethod
    def to_prepared_scene(coords_all, volumes_all, logits_all, dims_all) -> PreparedScene:
        """build_volumes' per-level results (COARSE -> FINE: voxel coordinates (n,3), 7-channel feature rows (n,7),
        matching logits (n,1) or None) -> a PreparedScene, without the reference's int64 tables / fp32 masks."""
        return PreparedScene.from_sparse(coords_all, volumes_all, logits_all, dims_all)
