"""Drop-in for the reference's ``FeatureNetwork`` (models/modules/feature_network.py:126-178): same constructor conf
(``d_in``, ``d_base``, ``d_out``), same parameter names (``encoder_layers.{i}.{0,1}.conv.weight``,
``decoder_layers.{i}.conv.weight``, ``out_layers.{i}.weight``), same outputs (coarse -> fine list of (nv, d_out, H/2^i,
W/2^i) tensors), computed by the hand-written kernels of ``csrc/fpn.cu``.  Inference only (outputs detached)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib

EPS = 1e-5       # nn.InstanceNorm2d default


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _ConvIN(nn.Module):
    """Conv2d / ConvTranspose2d (no bias) + InstanceNorm2d + ReLU: only the weight is a parameter (feature_network.py:6-25,
    56-75); the ``conv`` attribute keeps the reference's parameter path."""

    def __init__(self, c_in, c_out, stride, transposed=False):
        super().__init__()
        self.stride, self.transposed = stride, transposed
        if transposed:
            self.conv = nn.ConvTranspose2d(c_in, c_out, 3, stride=stride, padding=1, output_padding=1, bias=False)
        else:
            self.conv = nn.Conv2d(c_in, c_out, 3, stride=stride, padding=1, bias=False)


class _Raw:
    """A raw convolution output with the InstanceNorm statistics its consumers apply on load."""

    def __init__(self, x, stats):
        self.x, self.stats = x, stats


def _conv(x, stats, weight, stride, want_stats=True):
    lib = _lib.load()
    n, c_in, h, w = (int(v) for v in x.shape)
    c_out = int(weight.shape[0])
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    out = torch.empty((n, c_out, ho, wo), dtype=torch.float32, device=x.device)
    nblk = int(lib.surf_fpn_conv_blocks(ho, wo))
    sums = torch.empty((n * c_out, nblk, 2), dtype=torch.float64, device=x.device) if want_stats else None
    wt = weight.detach().to(torch.float32).contiguous()
    _lib.check(lib.surf_fpn_conv3x3(x.data_ptr(), stats.data_ptr() if stats is not None else None, wt.data_ptr(), n, c_in,
                                    h, w, c_out, stride, out.data_ptr(), sums.data_ptr() if want_stats else None,
                                    _stream()), "fpn_conv3x3")
    if not want_stats:
        return out
    st = torch.empty((n * c_out, 2), dtype=torch.float32, device=x.device)
    _lib.check(lib.surf_fpn_finish_stats(sums.data_ptr(), n * c_out, nblk, ho * wo, EPS, st.data_ptr(), _stream()),
               "fpn_finish_stats")
    return _Raw(out, st)


def _deconv(x, stats, weight):
    lib = _lib.load()
    n, c_in, h, w = (int(v) for v in x.shape)
    c_out = int(weight.shape[1])
    out = torch.empty((n, c_out, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
    nblk = int(lib.surf_fpn_deconv_blocks(2 * h, 2 * w))
    sums = torch.empty((n * c_out, nblk, 2), dtype=torch.float64, device=x.device)
    wt = weight.detach().to(torch.float32).contiguous()
    _lib.check(lib.surf_fpn_deconv3x3s2(x.data_ptr(), stats.data_ptr() if stats is not None else None, wt.data_ptr(), n,
                                        c_in, h, w, c_out, out.data_ptr(), sums.data_ptr(), _stream()), "fpn_deconv3x3s2")
    st = torch.empty((n * c_out, 2), dtype=torch.float32, device=x.device)
    _lib.check(lib.surf_fpn_finish_stats(sums.data_ptr(), n * c_out, nblk, 4 * h * w, EPS, st.data_ptr(), _stream()),
               "fpn_finish_stats")
    return _Raw(out, st)


def _norm_relu_add(a: _Raw, b: _Raw):
    lib = _lib.load()
    n, c, h, w = (int(v) for v in a.x.shape)
    out = torch.empty_like(a.x)
    _lib.check(lib.surf_fpn_norm_relu_add(a.x.data_ptr(), a.stats.data_ptr(), b.x.data_ptr() if b is not None else None,
                                          b.stats.data_ptr() if b is not None else None, n * c, h * w, out.data_ptr(),
                                          _stream()), "fpn_norm_relu_add")
    return out


class FeatureNetwork(nn.Module):
    def __init__(self, confs):
        super().__init__()
        d_in = confs.get_int("d_in")
        d_base = confs.get_int("d_base")
        d_outs = confs.get_list("d_out")            # fine to coarse
        self.num_stage = len(d_outs)
        self.encoder_layers = nn.ModuleList([])
        self.decoder_layers = nn.ModuleList([])
        self.out_layers = nn.ModuleList([])
        for i in range(self.num_stage):
            dim = d_base * 2 ** i
            self.encoder_layers.append(nn.Sequential(_ConvIN(d_in, dim, 2 if i > 0 else 1), _ConvIN(dim, dim, 1)))
            d_in = dim
            self.out_layers.append(nn.Conv2d(dim, d_outs[i], 3, 1, 1, bias=False))
            if i < self.num_stage - 1:
                self.decoder_layers.append(_ConvIN(d_base * 2 ** (i + 1), dim, 2, transposed=True))

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("surf_b200 FeatureNetwork needs CUDA tensors (there is no CPU fallback)")
        x = x.detach().to(torch.float32).contiguous()
        if x.shape[2] % (2 ** (self.num_stage - 1)) or x.shape[3] % (2 ** (self.num_stage - 1)):
            raise ValueError("image size must be a multiple of %d (the decoder doubles the coarser stage exactly)"
                             % 2 ** (self.num_stage - 1))
        with torch.cuda.device(x.device):
            e_outs = []
            cur, cur_stats = x, None
            for i in range(self.num_stage):
                a, b = self.encoder_layers[i][0], self.encoder_layers[i][1]
                r = _conv(cur, cur_stats, a.conv.weight, a.stride)
                r = _conv(r.x, r.stats, b.conv.weight, 1)
                e_outs.append(r)
                cur, cur_stats = r.x, r.stats
            # decoder (feature_network.py:164-168): d = relu(norm(deconv(prev))) + e_out[i]
            d_plain = [None] * self.num_stage
            prev_x, prev_stats = e_outs[-1].x, e_outs[-1].stats
            for i in range(self.num_stage - 2, -1, -1):
                up = _deconv(prev_x, prev_stats, self.decoder_layers[i].conv.weight)
                d_plain[i] = _norm_relu_add(up, e_outs[i])
                prev_x, prev_stats = d_plain[i], None
            outs = []
            for i in range(self.num_stage):
                if i == self.num_stage - 1:
                    outs.append(_conv(e_outs[i].x, e_outs[i].stats, self.out_layers[i].weight, 1, want_stats=False))
                else:
                    outs.append(_conv(d_plain[i], None, self.out_layers[i].weight, 1, want_stats=False))
        return outs[::-1]          # coarse to fine
