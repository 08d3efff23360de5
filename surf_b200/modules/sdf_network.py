"""SDFNetworkSparse — drop-in for models/modules/sdf_network.py:27-152.

Same constructor arguments, same geometric initialisation, same parameter names
(``lin{l}.weight_g / weight_v / bias``) so a reference checkpoint loads unchanged.  ``sdf`` and
``gradient`` run the hand-written sm_100a kernel (csrc/sdf_mlp.cu) through the C-ABI; there is no
PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..scene import AUX_SCENE_CACHE, PreparedScene
from .embedder import embedder_out_dim


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SDFNetworkSparse(nn.Module):
    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False, feat_channels=32, feat_multires=2):
        super().__init__()
        if feat_multires > 0:
            raise NotImplementedError("feat_multires > 0 is not used by any shipped conf (confs/*.conf: 0)")
        self.multires = int(multires)
        self.d_in_raw = int(d_in)
        d_in = embedder_out_dim(multires, d_in)
        dims = [d_in] + [d_hidden + feat_channels for _ in range(n_layers)] + [d_out]
        self.dims = dims
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        self.scale = float(scale)
        self.feat_channels = int(feat_channels)
        self.weight_norm = bool(weight_norm)
        for l in range(0, self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            if l < self.num_layers - 2:
                out_dim = out_dim - feat_channels
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:      # sdf_network.py:62-86: SDF ~ ||x|| - bias at initialisation
                if l == self.num_layers - 2:
                    sign = -1.0 if inside_outside else 1.0
                    torch.nn.init.normal_(lin.weight, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                    torch.nn.init.constant_(lin.bias, -sign * bias)
                    torch.nn.init.constant_(lin.weight[:, -feat_channels:], 0.0)
                    torch.nn.init.constant_(lin.bias[-feat_channels:], 0.0)
                elif multires > 0 and l == 0:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
                    torch.nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    torch.nn.init.constant_(lin.weight[:, -(dims[0] - 3 + feat_channels):], 0.0)
                else:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    torch.nn.init.constant_(lin.weight[:, -feat_channels:], 0.0)
            if weight_norm:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    lin = nn.utils.weight_norm(lin)     # parameter names weight_g / weight_v, as the reference
            setattr(self, "lin" + str(l), lin)
        self._net_owner = None      # set by ImplicitSurface: the object that builds the surf_net handle

    # -- C-ABI description of this network -----------------------------------------------------
    def fill_net_inputs(self, inp: "_lib.NetInputs", keep: list):
        n_lin = self.num_layers - 1
        if n_lin != _lib.SDF_LAYERS:
            raise NotImplementedError("the sm_100a SDF kernel is specialised for n_layers=6 (7 Linear layers)")
        if len(self.skip_in) > 1:
            raise NotImplementedError("at most one skip layer")
        inp.n_lin = n_lin
        inp.multires = self.multires
        inp.skip_layer = self.skip_in[0] if self.skip_in else -1
        inp.feat_channels = self.feat_channels
        inp.scale = self.scale
        for l in range(n_lin):
            lin = getattr(self, "lin" + str(l))
            if hasattr(lin, "weight_v"):
                v = lin.weight_v.detach().float().cpu().contiguous()
                g = lin.weight_g.detach().float().cpu().contiguous().reshape(-1)
                keep += [v, g]
                inp.h_weight_v[l], inp.h_weight_g[l] = v.data_ptr(), g.data_ptr()
            else:
                v = lin.weight.detach().float().cpu().contiguous()
                keep.append(v)
                inp.h_weight_v[l], inp.h_weight_g[l] = v.data_ptr(), None
            b = lin.bias.detach().float().cpu().contiguous()
            keep.append(b)
            inp.h_bias[l] = b.data_ptr()
            inp.out_dim[l], inp.in_dim[l] = int(v.shape[0]), int(v.shape[1])

    def _handles(self, volumes, indexes):
        if self._net_owner is None:
            raise RuntimeError("SDFNetworkSparse must be owned by an ImplicitSurface to build its device weights")
        owner = self._net_owner()
        net = owner.net_handle()
        scene = volumes if isinstance(volumes, PreparedScene) else AUX_SCENE_CACHE.get(volumes, indexes)
        return scene, net, int(owner.mlp_mode)

    # -- reference API -------------------------------------------------------------------------
    def sdf(self, x, volumes, indexes=None):
        """(n,3) -> (n,1)   (sdf_network.py:123-124).  ``volumes`` may be a PreparedScene."""
        scene, net, mode = self._handles(volumes, indexes)
        x = x.detach().to(torch.float32).contiguous()
        out = torch.empty((x.shape[0], 1), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().surf_sdf_points(scene.handle, net, x.data_ptr(), x.shape[0], out.data_ptr(), None,
                                                   mode, _stream()), "sdf_points")
        return out

    def gradient(self, x, volumes, indexes=None, with_sdf=False):
        """(gradients, smooth) like the reference (sdf_network.py:129-152): d sdf / d x (n,3) by the in-kernel analytic
        reverse pass of the render kernels, and smooth = d/dx sum_j (d sdf / d x_j) (the double-autograd term, :143-150)
        by the plain fp32 second-order kernel (surf_sdf_smooth).  ``with_sdf=True`` -> (sdf, gradients), no second-order
        pass."""
        scene, net, mode = self._handles(volumes, indexes)
        x = x.detach().to(torch.float32).contiguous()
        sdf = torch.empty((x.shape[0], 1), dtype=torch.float32, device=x.device)
        grad = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().surf_sdf_points(scene.handle, net, x.data_ptr(), x.shape[0], sdf.data_ptr(),
                                                   grad.data_ptr(), mode, _stream()), "sdf_points(grad)")
        if with_sdf:
            return sdf, grad
        return grad, self.smooth(x, scene)

    def smooth(self, x, volumes, indexes=None, flags=None, with_grad=False, mode=None):
        """Hessian(sdf) . (1,1,1) at x (n,3) -> (n,3).  ``flags`` (n,) uint8: bit 1 = evaluate, else 0 (the reference's
        masked-out default, implicit_surface.py:99).  ``with_grad``: also the first-order gradient of the same pass.
        ``mode``: kernel family (default: the module's ``mlp_mode``; MLP_FFMA = the plain fp32 kernel)."""
        scene, net, owner_mode = self._handles(volumes, indexes)
        x = x.detach().to(torch.float32).contiguous()
        sm = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
        gr = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device) if with_grad else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().surf_sdf_smooth(scene.handle, net, x.data_ptr(), x.shape[0],
                                                   flags.data_ptr() if flags is not None else None,
                                                   gr.data_ptr() if gr is not None else None, sm.data_ptr(),
                                                   int(owner_mode if mode is None else mode), _stream()),
                       "sdf_smooth")
        return (gr, sm) if with_grad else sm

    def forward(self, inputs, volumes, indexes=None):
        """(n,3) -> (n, d_out) = [sdf / scale, lin6 outputs 1..] exactly like the reference (sdf_network.py:95-121).
        The extra outputs are dead on the render path (implicit_surface.py:95-97 scatters them and never reads them),
        so ``sdf`` / ``gradient`` and the fused render kernels evaluate only row 0 of lin6; this entry point runs a plain
        fp32 kernel (surf_sdf_full)."""
        scene, net, _ = self._handles(volumes, indexes)
        x = inputs.detach().to(torch.float32).contiguous()
        d_out = int(self.dims[-1])
        out = torch.empty((x.shape[0], d_out), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().surf_sdf_full(scene.handle, net, x.data_ptr(), x.shape[0], out.data_ptr(), d_out,
                                                 _stream()), "sdf_full")
        return out
