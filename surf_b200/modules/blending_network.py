"""BlendingNetwork — drop-in for models/modules/blending_network.py:22-117 (IBRNet-style blending).

Same parameter names (``s``, ``ray_dir_fc.{0,2}``, ``base_fc.{0,2}``, ``vis_fc.{0,2}``,
``vis_fc2.{0,2}``, ``rgb_fc.{0,2,4}``); ``forward`` runs csrc/blend.cu through the C-ABI."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _lib

BLEND_LINEARS = ["ray_dir_fc.0", "ray_dir_fc.2", "base_fc.0", "base_fc.2", "vis_fc.0", "vis_fc.2",
                 "vis_fc2.0", "vis_fc2.2", "rgb_fc.0", "rgb_fc.2", "rgb_fc.4"]


def _weights_init(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class BlendingNetwork(nn.Module):
    def __init__(self, d_feature=16, anti_alias_pooling=True):
        super().__init__()
        if not anti_alias_pooling:
            raise NotImplementedError("anti_alias_pooling=False is not used by the reference confs")
        self.anti_alias_pooling = anti_alias_pooling
        self.d_feature = int(d_feature)
        self.s = nn.Parameter(torch.tensor(0.2), requires_grad=True)
        act = nn.ELU(inplace=True)
        c = d_feature + 3
        self.ray_dir_fc = nn.Sequential(nn.Linear(4, 16), act, nn.Linear(16, c), act)
        self.base_fc = nn.Sequential(nn.Linear(c * 3, 64), act, nn.Linear(64, 32), act)
        self.vis_fc = nn.Sequential(nn.Linear(32, 32), act, nn.Linear(32, 33), act)
        self.vis_fc2 = nn.Sequential(nn.Linear(32, 32), act, nn.Linear(32, 1), nn.Sigmoid())
        self.rgb_fc = nn.Sequential(nn.Linear(32 + 1 + 4, 16), act, nn.Linear(16, 8), act, nn.Linear(8, 1))
        self.base_fc.apply(_weights_init)
        self.vis_fc2.apply(_weights_init)
        self.vis_fc.apply(_weights_init)
        self.rgb_fc.apply(_weights_init)
        self._net_owner = None

    def fill_net_inputs(self, inp: "_lib.NetInputs", keep: list):
        if self.d_feature != 16:
            raise NotImplementedError("the sm_100a blending kernel is specialised for d_feature=16")
        inp.d_feature = self.d_feature
        inp.blend_s = float(self.s.detach().cpu())
        mods = dict(self.named_modules())
        for i, name in enumerate(BLEND_LINEARS):
            lin = mods[name]
            w = lin.weight.detach().float().cpu().contiguous()
            b = lin.bias.detach().float().cpu().contiguous()
            keep += [w, b]
            inp.h_blend_w[i], inp.h_blend_b[i] = w.data_ptr(), b.data_ptr()

    def forward(self, rgb_feat, ray_diff, mask):
        """rgb_feat (n,V,19), ray_diff (n,V,4), mask (n,V) bool -> rgb (n,3)."""
        if self._net_owner is None:
            raise RuntimeError("BlendingNetwork must be owned by an ImplicitSurface to build its device weights")
        owner = self._net_owner()
        net = owner.net_handle()
        n, V, c = rgb_feat.shape
        if c != self.d_feature + 3:
            raise ValueError("expected %d channels" % (self.d_feature + 3))
        f = rgb_feat.detach().to(torch.float32).contiguous()
        r = ray_diff.detach().to(torch.float32).contiguous()
        m = mask.detach().to(torch.uint8).contiguous()
        out = torch.empty((n, 3), dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device):
            _lib.check(_lib.load().surf_blend(net, f.data_ptr(), r.data_ptr(), m.data_ptr(), n, V, out.data_ptr(),
                                              int(owner.mlp_mode),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "blend")
        return out
