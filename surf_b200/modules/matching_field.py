"""MatchingField — drop-in for models/modules/matching_field.py:8-141 (the upstream user of the
sampler's probe kernel).

``forward(ipts, matching_volume, stage_idx, range_ratios, pre_depths=None, perturb=False)`` renders a depth map per view
from the dense matching volume (one CUDA kernel per view + the bilinear up-sampling to the image size) and returns
``(render_depths, occ_regs)`` like the reference.  The matching volume is converted once (cached on tensor identity);
the per-view camera algebra (3x3 inverses) is done on the host with the reference's own torch ops.  Outputs are
detached (the reference back-propagates through the probe into the cost-volume network during training; autograd
through the kernels is not provided).  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from ..scene import AUX_SCENE_CACHE, PreparedScene


def _stream():
    return torch.cuda.current_stream().cuda_stream


class MatchingField(nn.Module):
    def __init__(self, confs):
        super().__init__()
        self.n_samples_depths = confs.get_list("n_samples_depths")
        self.n_importance_depths = confs.get_list("n_importance_depths")
        self.up_sample_steps = confs.get_list("up_sample_steps")
        self.depth_res_levels = confs.get_list("depth_res_levels")

    @staticmethod
    def _scene(matching_volume):
        if isinstance(matching_volume, PreparedScene):
            return matching_volume
        return AUX_SCENE_CACHE.get([], [], None, matching_volume)

    def depth_map(self, scene, intrs, c2ws, view, near, far, img_hw, stage_idx, range_ratios, pre_depth=None, t_rand=None):
        """One view: -> (depth (img_h,img_w), occ_reg scalar tensor, depth at the stage's resolution (h,w))."""
        lib = _lib.load()
        dev = scene.device
        img_h, img_w = int(img_hw[0]), int(img_hw[1])
        level = int(self.depth_res_levels[stage_idx])
        h, w = img_h // level, img_w // level
        n = int(self.n_samples_depths[stage_idx])
        K = intrs.detach().float().cpu()
        c2w = c2ws.detach().float().cpu()
        p = _lib.DepthMapParams()
        p.Kinv[:] = K.inverse()[view, :3, :3].reshape(-1).tolist()
        p.R[:] = c2w[view, :3, :3].reshape(-1).tolist()
        p.C[:] = c2w[view, :3, 3].reshape(-1).tolist()
        p.Rinv2[:] = torch.inverse(c2w[view, None, :3, :3])[0, 2].reshape(-1).tolist()
        p.near, p.far = float(near), float(far)
        two = pre_depth is not None
        p.n_windows = 2 if two else 1
        p.n_samples = n
        if two:
            # fp32 like the reference's tensor arithmetic (range * ratio with a python float)
            p.ratio[0] = float(range_ratios[stage_idx])
            p.ratio[1] = float(range_ratios[stage_idx - 1])
        p.h, p.w, p.img_h, p.img_w = h, w, img_h, img_w
        f32 = dict(dtype=torch.float32, device=dev)
        lin = torch.linspace(0.0, 1.0, n).to(dev)
        tx = torch.linspace(0, img_w - 1, w).to(dev)
        ty = torch.linspace(0, img_h - 1, h).to(dev)
        low = torch.empty((h, w), **f32)
        full = torch.empty((img_h, img_w), **f32)
        occ = torch.empty((h * w, 3), **f32)
        pre = pre_depth.detach().to(**f32).contiguous() if two else None
        tr = t_rand.to(**f32).contiguous() if t_rand is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.surf_depth_map(scene.handle, C.byref(p), lin.data_ptr(), tx.data_ptr(), ty.data_ptr(),
                                          pre.data_ptr() if pre is not None else None,
                                          tr.data_ptr() if tr is not None else None, low.data_ptr(), occ.data_ptr(),
                                          full.data_ptr(), _stream()), "depth_map")
        s = occ.sum(dim=0)
        occ_reg = s[0] / (6.0 * h * w) + s[1] / (s[2] + 1e-10)
        return full, occ_reg, low

    def forward(self, ipts, matching_volume, stage_idx, range_ratios, pre_depths=None, perturb=False):
        near_fars, c2ws, intrs = ipts["near_fars"], ipts["c2ws"], ipts["intrs"]
        src_idx = ipts["src_idx"] if "src_idx" in ipts else 0
        img_hw = ipts["imgs"].shape[-2:]
        scene = self._scene(matching_volume)
        level = int(self.depth_res_levels[stage_idx])
        n_rays = (img_hw[0] // level) * (img_hw[1] // level)
        nf = near_fars.detach().float().cpu()
        depths, occs = [], []
        for i in range(intrs.shape[0]):
            t_rand = None
            if perturb and (i == 0 or i == src_idx):
                # the reference's stream: one rand([B,1]) per window from the global CPU generator (:34)
                nw = 1 if pre_depths is None else 2
                t_rand = torch.cat([torch.rand([n_rays, 1]) for _ in range(nw)], dim=1)
            d, occ, _ = self.depth_map(scene, intrs, c2ws, i, nf[i, 0], nf[i, 1], img_hw, stage_idx, range_ratios,
                                       None if pre_depths is None else pre_depths[i], t_rand)
            depths.append(d)
            occs.append(occ)
        return depths, occs
