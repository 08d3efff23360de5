"""CPU oracle: a restatement of SuRF's per-ray volume-rendering hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker / the CPU
baseline.  ``surf_b200/`` never imports it and has no CPU fallback.

What it is: the algorithm of the reference's hot path (SURVEY.md §8a rows
A1-A11) restated as plain functions over fp32 torch CPU tensors, each citing
the reference file:line it follows.  It deliberately uses the same ATen
primitives as the reference wherever a result is integer-sensitive
(``F.grid_sample`` nearest / ``torch.linspace`` / ``torch.sort`` /
``torch.inverse``), so on the same torch build it reproduces the reference
bit-for-bit for masks and indices.

Parity pin: ``oracle/make_golden.py`` imports the UNMODIFIED reference from
``/root/reference`` (three import shims, ``oracle/ref_loader.py``), runs it on
small synthetic scenes and stores inputs+outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors
(integers bit-exact, floats <= 1e-5).  The reference ships no tests or golden
vectors of its own.

Not restated (SURVEY.md §8f "next", F1): the training-only extras
``smooth`` (second-order autograd, sdf_network.py:143-150) and
``surface_patch_warp2`` (projector.py:560-645).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
class OracleNet:
    """Plain-tensor view of an ``ImplicitSurface`` state_dict (reference param names).

    ``sd`` keys: ``sdf_network.lin{l}.{weight_g,weight_v,bias}`` (or ``.weight`` when not
    weight-normed), ``color_network.*``, ``deviation_network.variance``.
    """

    def __init__(self, sd: Dict[str, torch.Tensor], n_samples=(64, 32, 24, 16),
                 sample_ranges=(1.0, 0.4, 0.1, 0.01), n_depth=256, perturb=1.0,
                 multires=4, skip_in=(3,), scale=1.0, n_layers=6):
        sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items()}
        self.n_samples = [int(x) for x in n_samples]
        self.sample_ranges = [float(x) for x in sample_ranges]
        self.n_depth = int(n_depth)
        self.perturb = float(perturb)
        self.multires = int(multires)
        self.skip_in = tuple(int(s) for s in skip_in)
        self.scale = float(scale)
        self.num_lin = n_layers + 1
        self.W: List[torch.Tensor] = []
        self.b: List[torch.Tensor] = []
        for l in range(self.num_lin):
            p = "sdf_network.lin%d." % l
            if p + "weight_g" in sd:
                # nn.utils.weight_norm(dim=0): W = g * v / ||v||_row   (sdf_network.py:88-89)
                W = torch._weight_norm(sd[p + "weight_v"], sd[p + "weight_g"], 0)
            else:
                W = sd[p + "weight"]
            self.W.append(W.contiguous())
            self.b.append(sd[p + "bias"].contiguous())
        c = "color_network."
        self.color = {k[len(c):]: v for k, v in sd.items() if k.startswith(c)}
        self.variance = sd["deviation_network.variance"].reshape(())


# ----------------------------------------------------------------------------------------------
# dense volume lookups  (projector.py:392-420)
# ----------------------------------------------------------------------------------------------
def lookup_dense(pts: torch.Tensor, volumes, mode: str) -> torch.Tensor:
    """``F.grid_sample`` of (1,C,N,N,N) volumes at ``pts`` (n,3) -> (n, sum C).

    Grid = pts flipped so world (x,y,z) <-> volume axes (D,H,W); default align_corners=False and
    zero padding (projector.py:398,406,415; quirk Q4)."""
    pts = pts.reshape(-1, 3)
    n = pts.shape[0]
    grid = pts.flip(-1)[None, None, None]
    vols = [volumes] if isinstance(volumes, torch.Tensor) else list(volumes)
    cols = []
    for v in vols:
        s = F.grid_sample(v, grid, mode=mode, align_corners=False)
        cols.append(s.reshape(-1, n).t().contiguous())
    return torch.cat(cols, dim=-1)


def point_mask(pts: torch.Tensor, mask_volumes) -> torch.Tensor:
    """voxel mask = nearest lookup in every level's 0/1 volume, OR-ed (implicit_surface.py:86; Q5)."""
    return lookup_dense(pts, mask_volumes, "nearest").any(dim=-1)


# ----------------------------------------------------------------------------------------------
# sparse trilinear lookup  (projector.py:217-390; quirk Q13)
# ----------------------------------------------------------------------------------------------
def sparse_trilinear(volume: torch.Tensor, index: torch.Tensor, pts_zyx: torch.Tensor) -> torch.Tensor:
    """volume (nvox,c), index (N,N,N) int64 (-1 empty), pts already flipped to (z,y,x) -> (n,c).

    Differentiable w.r.t. ``pts_zyx`` through the interpolation weights only (corner indices are
    integers), exactly like projector.py:238-283."""
    n = pts_zyx.shape[0]
    c = volume.shape[1]
    dims = torch.tensor(list(index.shape), dtype=pts_zyx.dtype)
    voxel = (torch.ones(3, dtype=pts_zyx.dtype) - (-torch.ones(3, dtype=pts_zyx.dtype))) / (dims - 1)
    coords = (pts_zyx - (-torch.ones(3, dtype=pts_zyx.dtype))[None]) / voxel[None]     # :231-232
    cx, cy, cz = coords[:, 0], coords[:, 1], coords[:, 2]
    with torch.no_grad():
        x0 = torch.floor(cx).long()
        y0 = torch.floor(cy).long()
        z0 = torch.floor(cz).long()
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    IW, IH, ID = index.shape
    flat = index.reshape(-1)
    out = None
    # corner order and weight expressions follow projector.py:274-283 / 360-371 (bnw ... fse)
    for zz, wz_hi in ((z0, True), (z1, False)):
        for yy, wy_hi in ((y0, True), (y1, False)):
            for xx, wx_hi in ((x0, True), (x1, False)):
                wx = (x1 - cx) if wx_hi else (cx - x0)
                wy = (y1 - cy) if wy_hi else (cy - y0)
                wz = (z1 - cz) if wz_hi else (cz - z0)
                w = wx * wy * wz
                with torch.no_grad():
                    lin = (zz.clamp(0, ID - 1) * ID ** 2 + yy.clamp(0, IH - 1) * IW + xx.clamp(0, IW - 1))
                    row = flat[lin]                                                   # :323-340
                    ok = row != -1
                val = torch.zeros(n, c, dtype=volume.dtype)
                val[ok] = volume[row[ok]]                                             # :342-358
                term = val * w[:, None]
                out = term if out is None else out + term
    return out


def voxel_round_distance(pts: torch.Tensor, mask_volumes) -> torch.Tensor:
    """World-space distance of each point to the nearest decision boundary of the 'nearest' mask lookup, min over
    levels and axes.  F.grid_sample(nearest, align_corners=False) picks voxel nearbyint(u), u = ((c+1) N - 1)/2
    (ATen GridSampler.h), so the decision (incl. the in/out-of-bounds one at u = -0.5 and N - 0.5) flips where
    u = k + 0.5; du/dc = N/2.  Test helper: two evaluations whose sample positions differ by rounding noise may
    legitimately disagree on the mask only for points closer to such a boundary than that noise."""
    vols = [mask_volumes] if isinstance(mask_volumes, torch.Tensor) else list(mask_volumes)
    p = pts.reshape(-1, 3).double()
    dist = torch.full((p.shape[0],), float("inf"), dtype=torch.float64)
    for v in vols:
        n = float(v.shape[-1])
        u = ((p + 1.0) * n - 1.0) * 0.5
        d = (u - torch.floor(u) - 0.5).abs() * (2.0 / n)
        dist = torch.minimum(dist, d.min(dim=1)[0])
    return dist


def voxel_face_distance(pts: torch.Tensor, indexes: Sequence[torch.Tensor]) -> torch.Tensor:
    """min over levels and axes of |c - round(c)| / ulp(c) for the sparse-grid coordinate c = (p+1)/voxel
    (projector.py:231-242).  The trilinear feature lookup is continuous across voxel faces but its
    GRADIENT is not: a point within rounding noise of a face (c is an exact integer in fp32 for about
    4e-6 of all points per axis and level) gets a different d sdf / d x when its position changes by one
    ulp.  Test helper: marks samples whose gradient the reference itself does not define robustly."""
    p = pts.flip(-1)
    dist = torch.full((pts.shape[0],), float("inf"))
    for idx in indexes:
        n = idx.shape[0]
        voxel = torch.tensor(2.0) / (torch.tensor(float(n)) - 1)
        c = (p + 1.0) / voxel
        # in units of the fp32 spacing at c, so that "< 1.5" means "within rounding noise of a face"
        ulp = torch.finfo(torch.float32).eps * c.abs().clamp(min=1.0)
        dist = torch.minimum(dist, ((c - torch.round(c)).abs() / ulp).min(dim=1)[0])
    return dist


def voxel_face_margin(pts: torch.Tensor, indexes: Sequence[torch.Tensor], slack: torch.Tensor) -> torch.Tensor:
    """voxel_face_distance with a per-point position uncertainty ``slack`` (world units, e.g. the measured difference
    of the two paths' sample depths): min over levels and axes of (|c - round(c)| - slack / voxel) / ulp(c).  A value
    below ~2.5 means the two paths may legitimately have evaluated the point on different sides of a voxel face."""
    p = pts.flip(-1).double()
    out = torch.full((pts.shape[0],), float("inf"), dtype=torch.float64)
    for idx in indexes:
        n = idx.shape[0]
        voxel = float(torch.tensor(2.0) / (torch.tensor(float(n)) - 1))
        c = (p + 1.0) / voxel
        ulp = torch.finfo(torch.float32).eps * c.abs().clamp(min=1.0)
        out = torch.minimum(out, (((c - torch.round(c)).abs() - slack.double()[:, None] / voxel) / ulp).min(dim=1)[0])
    return out


def lookup_sparse(pts: torch.Tensor, volumes: Sequence[torch.Tensor], indexes: Sequence[torch.Tensor]):
    """(n,3) -> (n, 7*levels); levels concatenated in the order given (fine->coarse) (:377-390)."""
    p = pts.flip(-1)
    return torch.cat([sparse_trilinear(v, i, p) for v, i in zip(volumes, indexes)], dim=-1)


# ----------------------------------------------------------------------------------------------
# SDF MLP  (sdf_network.py:95-152, embedder.py:6-51)
# ----------------------------------------------------------------------------------------------
def positional_encoding(x: torch.Tensor, multires: int) -> torch.Tensor:
    if multires <= 0:
        return x
    outs = [x]
    for f in 2.0 ** torch.linspace(0.0, multires - 1, multires):
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, dim=-1)


def softplus100(x):
    return F.softplus(x, beta=100)


def sdf_forward(net: OracleNet, pts: torch.Tensor, volumes, indexes, feats: Optional[torch.Tensor] = None):
    """Full (n,129) output of SDFNetworkSparse.forward (sdf_network.py:95-121)."""
    if feats is None:
        feats = lookup_sparse(pts.clone(), volumes, indexes)
    pe = positional_encoding(pts * net.scale, net.multires)
    h = pe
    last = net.num_lin - 1
    for l in range(net.num_lin):
        if l in net.skip_in:
            h = torch.cat([h, pe], dim=-1) / SQRT2
        if 0 < l:
            h = torch.cat([h, feats], dim=-1)
        h = F.linear(h, net.W[l], net.b[l])
        if l < last:
            h = softplus100(h)
    return torch.cat([h[:, :1] / net.scale, h[:, 1:]], dim=-1)


def sdf_only(net: OracleNet, pts, volumes, indexes):
    return sdf_forward(net, pts, volumes, indexes)[:, :1]


def sdf_gradient(net: OracleNet, pts: torch.Tensor, volumes, indexes):
    """d sdf / d x by autograd, like sdf_network.py:129-141 (first-order part only).

    Returns (sdf (n,1), grad (n,3)), both detached."""
    with torch.enable_grad():
        x = pts.detach().clone().requires_grad_(True)
        y = sdf_only(net, x, volumes, indexes)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y))
    return y.detach(), g.detach()


def sdf_gradient_analytic(net: OracleNet, pts: torch.Tensor, volumes, indexes):
    """The same gradient by an explicit reverse pass (the derivation the CUDA kernel implements).

    dsdf/dx = J_pe^T g_pe + J_feat^T g_feat where g_pe collects the input-gradients of lin0 and of
    the skip layer's PE columns, and g_feat those of the 28 feature columns of lin1..lin6."""
    n = pts.shape[0]
    p = pts.flip(-1)
    # feature values and their Jacobian w.r.t. the *flipped* point, level by level
    feats, jac = [], []
    for vol, idx in zip(volumes, indexes):
        N = idx.shape[0]
        voxel = torch.tensor(2.0, dtype=torch.float32) / (torch.tensor(float(N)) - 1)
        c = (p + 1.0) / voxel
        i0 = torch.floor(c)
        fr1 = c - i0                 # (coord - corner0)
        fr0 = (i0 + 1) - c           # (corner1 - coord)
        i0 = i0.long()
        f = torch.zeros(n, vol.shape[1])
        J = torch.zeros(n, vol.shape[1], 3)
        flat = idx.reshape(-1)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    wx = fr1[:, 0] if dx else fr0[:, 0]
                    wy = fr1[:, 1] if dy else fr0[:, 1]
                    wz = fr1[:, 2] if dz else fr0[:, 2]
                    sx, sy, sz = (1.0 if dx else -1.0), (1.0 if dy else -1.0), (1.0 if dz else -1.0)
                    lin = ((i0[:, 2] + dz).clamp(0, N - 1) * N * N + (i0[:, 1] + dy).clamp(0, N - 1) * N
                           + (i0[:, 0] + dx).clamp(0, N - 1))
                    row = flat[lin]
                    ok = row != -1
                    val = torch.zeros(n, vol.shape[1])
                    val[ok] = vol[row[ok]]
                    f += val * (wx * wy * wz)[:, None]
                    J[:, :, 0] += val * (sx * wy * wz / voxel)[:, None]
                    J[:, :, 1] += val * (wx * sy * wz / voxel)[:, None]
                    J[:, :, 2] += val * (wx * wy * sz / voxel)[:, None]
        feats.append(f)
        jac.append(J)
    feats = torch.cat(feats, dim=1)
    jac = torch.cat(jac, dim=1)                      # (n, 28, 3) in flipped (z,y,x) order
    x = pts * net.scale
    pe = positional_encoding(x, net.multires)
    last = net.num_lin - 1
    h = pe
    sig = []
    nfeat = feats.shape[1]
    for l in range(net.num_lin):
        if l in net.skip_in:
            h = torch.cat([h, pe], dim=-1) / SQRT2
        if l > 0:
            h = torch.cat([h, feats], dim=-1)
        z = F.linear(h, net.W[l], net.b[l])
        if l < last:
            sig.append(torch.sigmoid(100.0 * z))     # softplus'(z), beta=100
            h = softplus100(z)
        else:
            h = z
    sdf = h[:, :1] / net.scale
    # reverse pass
    g_pe = torch.zeros(n, pe.shape[1])
    g_feat = torch.zeros(n, nfeat)
    delta = torch.zeros(n, net.W[last].shape[0])
    delta[:, 0] = 1.0 / net.scale
    for l in range(last, -1, -1):
        if l < last:
            delta = delta * sig[l]
        gin = delta @ net.W[l]                       # gradient w.r.t. this layer's input
        if l > 0:
            g_feat += gin[:, -nfeat:]
            gin = gin[:, :-nfeat]
        if l in net.skip_in:
            gin = gin / SQRT2
            g_pe += gin[:, -pe.shape[1]:]
            gin = gin[:, :-pe.shape[1]]
        if l == 0:
            g_pe += gin
        delta = gin
    # d PE / d x
    gx = g_pe[:, 0:3].clone()
    k = 3
    for f in 2.0 ** torch.linspace(0.0, net.multires - 1, net.multires):
        gx += g_pe[:, k:k + 3] * torch.cos(x * f) * f
        gx -= g_pe[:, k + 3:k + 6] * torch.sin(x * f) * f
        k += 6
    gx = gx * net.scale
    gp = torch.einsum("nc,ncd->nd", g_feat, jac)     # w.r.t. flipped point
    return sdf, gx + gp.flip(-1)


# ----------------------------------------------------------------------------------------------
# multi-view projection + gather  (projector.py:485-556; quirk Q14)
# ----------------------------------------------------------------------------------------------
def ray_direction_diff(pts, ref_c2w, src_c2ws):
    """(n,3) -> (n,V,4): [normalised difference of unit vectors to ref/src centres, their dot]."""
    to_ref = ref_c2w[None, :3, 3] - pts                                  # (n,3)
    to_ref = to_ref / (torch.norm(to_ref, dim=-1, keepdim=True) + 1e-6)
    to_src = src_c2ws[:, None, :3, 3] - pts[None]                        # (V,n,3)
    to_src = to_src / (torch.norm(to_src, dim=-1, keepdim=True) + 1e-6)
    diff = to_ref[None] - to_src
    dot = (to_ref[None] * to_src).sum(-1, keepdim=True)
    direction = diff / torch.clamp(torch.norm(diff, dim=-1, keepdim=True), min=1e-6)
    return torch.cat([direction, dot], dim=-1).permute(1, 0, 2).contiguous()


def projection_border_distance(pts, intrs, c2ws, features):
    """min over source views and pyramid levels of the distance (in that level's pixels) between a
    point's projection and the nearest image border / the w=0 plane (in camera units).  The per-view
    validity mask (projector.py:536) flips across those borders, so a point closer than rounding noise
    to one of them has no well-defined mask (e.g. every ray of pixel row 0 projects to y=0 in a source
    view that differs from the reference only by an x offset).  Test helper."""
    src_K, src_c2w = intrs[1:], c2ws[1:]
    n = pts.shape[0]
    homog = torch.cat([pts.t().contiguous(), torch.ones(1, n)], dim=0)
    cam = torch.matmul(torch.inverse(src_c2w), homog[None])[:, :3]
    dist = torch.full((n,), float("inf"))
    for i, feat in enumerate(features):
        K = src_K.clone()
        K[:, :2] = K[:, :2] * (0.5 ** i)
        h, w = feat.shape[-2:]
        uvw = torch.matmul(K[:, :3, :3], cam)
        xy = uvw[:, :2] / uvw[:, 2:]
        d = torch.stack([xy[:, 0].abs(), (xy[:, 0] - w).abs(), xy[:, 1].abs(), (xy[:, 1] - h).abs(),
                         uvw[:, 2].abs()], dim=0).min(dim=0)[0]          # (V,n)
        dist = torch.minimum(dist, d.min(dim=0)[0])
    return dist


def lookup_feature(pts, imgs, intrs, c2ws, features):
    """-> feat_views (n,V,19) = [rgb3, f0(4), f1(4), f2(4), f3(4)], ray_diff (n,V,4), mask (n,V) bool."""
    src_K, src_c2w, ref_c2w = intrs[1:], c2ws[1:], c2ws[0]
    V = src_K.shape[0]
    n = pts.shape[0]
    ray_diff = ray_direction_diff(pts, ref_c2w, src_c2w)
    homog = torch.cat([pts.t().contiguous(), torch.ones(1, n)], dim=0)    # (4,n)
    cam = torch.matmul(torch.inverse(src_c2w), homog[None])[:, :3]        # (V,3,n)   :529
    feats_per_level = []
    mask_all = None
    rgb = None
    for i, feat in enumerate(features):
        K = src_K.clone()
        K[:, :2] = K[:, :2] * (0.5 ** i)                                  # :525
        h, w = feat.shape[-2:]
        uvw = torch.matmul(K[:, :3, :3], cam)                             # :530
        xy = uvw[:, :2] / uvw[:, 2:]                                      # no guard on w (:531)
        gx = xy[:, 0] / ((w - 1) / 2) - 1
        gy = xy[:, 1] / ((h - 1) / 2) - 1
        m = (uvw[:, 2] > 0) & (xy[:, 0] >= 0) & (xy[:, 0] < w) & (xy[:, 1] >= 0) & (xy[:, 1] < h)
        m = m.t()
        mask_all = m if mask_all is None else (mask_all & m)
        grid = torch.stack([gx, gy], dim=-1)[:, :, None]                  # (V,n,1,2)
        s = F.grid_sample(feat[1:], grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        feats_per_level.append(s.reshape(V, feat.shape[1], n).permute(2, 0, 1))
        if i == 0:
            s = F.grid_sample(imgs[1:], grid, mode="bilinear", padding_mode="zeros", align_corners=False)
            rgb = s.reshape(V, 3, n).permute(2, 0, 1)
    return torch.cat([rgb] + feats_per_level, dim=2).contiguous(), ray_diff, mask_all.contiguous()


# ----------------------------------------------------------------------------------------------
# colour blending MLP  (blending_network.py:69-117)
# ----------------------------------------------------------------------------------------------
def _seq(x, cw, name, idxs, final_act=True):
    for j, i in enumerate(idxs):
        x = F.linear(x, cw["%s.%d.weight" % (name, i)], cw["%s.%d.bias" % (name, i)])
        if final_act or j + 1 < len(idxs):
            x = F.elu(x)
    return x


def blend(net: OracleNet, feat_views, ray_diff, mask, e_ulp=None):
    """``e_ulp`` (n,V) integer: test-only knob that moves the pooling exponentials by that many fp32
    ulps, used by tests/helpers.blend_envelope to measure how ill-conditioned the reference's
    anti-alias weights are at a point (they subtract nearly equal exponentials)."""
    cw = net.color
    m = mask[:, :, None].to(feat_views.dtype)
    V = feat_views.shape[1]
    rgb_in = feat_views[..., :3]
    x = feat_views + _seq(ray_diff, cw, "ray_dir_fc", (0, 2))
    dot = ray_diff[..., 3:4]
    e = torch.exp(torch.abs(cw["s"]) * (dot - 1))
    if e_ulp is not None:
        up = torch.nextafter(e, torch.full_like(e, 4.0))
        dn = torch.nextafter(e, torch.full_like(e, -4.0))
        sh = e_ulp[:, :, None].to(e.dtype)          # any integer number of ulps
        e = torch.where(sh > 0, e + sh * (up - e), torch.where(sh < 0, e + sh * (e - dn), e))
    wgt = (e - torch.min(e, dim=1, keepdim=True)[0]) * m
    wgt = wgt / (torch.sum(wgt, dim=1, keepdim=True) + 1e-8)
    mean = torch.sum(x * wgt, dim=1, keepdim=True)
    var = torch.sum(wgt * (x - mean) ** 2, dim=1, keepdim=True)
    g = torch.cat([mean, var], dim=-1).expand(-1, V, -1)
    y = _seq(torch.cat([g, x], dim=-1), cw, "base_fc", (0, 2))
    yv = _seq(y * wgt, cw, "vis_fc", (0, 2))
    res, vis = yv[..., :-1], yv[..., -1:]
    vis = torch.sigmoid(vis) * m
    y = y + res
    v2 = F.linear(F.elu(F.linear(y * vis, cw["vis_fc2.0.weight"], cw["vis_fc2.0.bias"])),
                  cw["vis_fc2.2.weight"], cw["vis_fc2.2.bias"])
    vis = torch.sigmoid(v2) * m
    z = torch.cat([y, vis, ray_diff], dim=-1)
    z = F.elu(F.linear(z, cw["rgb_fc.0.weight"], cw["rgb_fc.0.bias"]))
    z = F.elu(F.linear(z, cw["rgb_fc.2.weight"], cw["rgb_fc.2.bias"]))
    z = F.linear(z, cw["rgb_fc.4.weight"], cw["rgb_fc.4.bias"])
    z = z.masked_fill(m == 0, -1e9)
    bw = F.softmax(z, dim=1)
    return torch.sum(rgb_in * bw, dim=1)


# ----------------------------------------------------------------------------------------------
# ray sampling  (implicit_surface.py:268-311; quirks Q1-Q3)
# ----------------------------------------------------------------------------------------------
def draw_t_rand(batch, n_stages=4):
    """Draws the per-stage jitter from torch's global CPU generator in the reference's order:
    one rand([B,1]) per stage (implicit_surface.py:276,305).  Returns (B, n_stages) in [0,1)."""
    return torch.cat([torch.rand([batch, 1]) for _ in range(n_stages)], dim=1)


def sample_z(net: OracleNet, rays_o, rays_d, near, far, matching_volume, t_rand: Optional[torch.Tensor]):
    """-> sorted z_vals (B, sum n_samples), surf_z (B,1).  ``t_rand`` (B,4) in [0,1) (raw rand, the
    0.5 shift is applied here); None or perturb<=0 -> no jitter."""
    B = rays_o.shape[0]
    n0 = net.n_samples[0]
    jitter = net.perturb > 0 and t_rand is not None
    lin = torch.linspace(0.0, 1.0, n0)
    z0 = near + (far - near) * lin[None, :]
    if jitter:
        z0 = z0 + (t_rand[:, 0:1] - 0.5) * 2.0 / n0
    stages = [z0]
    span = far - near
    lin = torch.linspace(0.0, 1.0, net.n_depth)
    zp = near + (far - near) * lin[None, :]
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * zp[..., :, None]).reshape(-1, 3)
    logit = lookup_dense(pts, matching_volume, "bilinear").reshape(B, -1)
    w = F.softmax(logit, dim=-1)
    surf_z = (zp * w).sum(dim=1, keepdim=True)
    for s, (ratio, n) in enumerate(zip(net.sample_ranges[1:], net.n_samples[1:]), start=1):
        lo = surf_z - span * ratio
        hi = surf_z + span * ratio
        lo = torch.where(hi > far, lo - (hi - far), lo)
        hi = torch.where(lo < near, hi + (near - lo), hi)
        lo = torch.clamp(lo, near, far)
        hi = torch.clamp(hi, near, far)
        lin = torch.linspace(0.0, 1.0, n)
        zs = lo + (hi - lo) * lin[None, :]
        if jitter:
            zs = zs + (t_rand[:, s:s + 1] - 0.5) * (hi - lo) / n
        stages.append(zs)
    z, _ = torch.sort(torch.cat(stages, dim=-1), dim=-1)
    return z, surf_z


# ----------------------------------------------------------------------------------------------
# compositing  (implicit_surface.py:126-216; quirks Q8-Q12)
# ----------------------------------------------------------------------------------------------
def composite(rays_o, rays_d, mid_z, dists, vmask, sdf, grad, color, inv_s, rot0_inv, cos_anneal_ratio=1.0,
              prev_idx=None):
    """NeuS alpha, transmittance, colour / normal / depth compositing and the first zero-crossing depth from
    per-point values.  vmask (B,S) float 0/1; sdf (P,1), grad (P,3), color (P,3) with the masked-out defaults
    already in place (Q7).  dtype-generic (the tests run it in float64 with autograd to get the sensitivity of
    every output to the per-point values).  ``prev_idx`` (B,1) pins the crossing index (discrete) when given."""
    B, S = mid_z.shape
    dt = sdf.dtype
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, S, 3).reshape(-1, 3)
    true_cos = (dirs * grad).sum(-1, keepdim=True)
    r = cos_anneal_ratio
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - r) + F.relu(-true_cos) * r)
    iter_cos = iter_cos * vmask.reshape(-1, 1)
    step = iter_cos.clip(-10.0, 10.0) * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid((sdf - step) * inv_s)
    next_cdf = torch.sigmoid((sdf + step) * inv_s)
    alpha = (((prev_cdf - next_cdf) + 1e-5) / (prev_cdf + 1e-5)).reshape(B, S).clip(0.0, 1.0)
    alpha = alpha * vmask

    pnorm = torch.linalg.norm(pts, ord=2, dim=-1).reshape(B, S)
    inside = (pnorm < 1.0).to(dt) * vmask
    relax_inside = (pnorm < 1.2).to(dt) * vmask

    trans = torch.cumprod(torch.cat([torch.ones(B, 1, dtype=dt), 1.0 - alpha + 1e-7], dim=-1), dim=-1)[:, :-1]
    weights = alpha * trans
    weight_sum = weights.sum(dim=-1, keepdim=True)
    color_fine = (color.reshape(B, S, 3) * weights[:, :, None]).sum(dim=1)
    g3 = grad.reshape(B, S, 3)
    rot = rot0_inv.to(dt)
    normal = torch.matmul(rot[None], (g3 * weights[:, :, None]).sum(dim=1)[:, :, None]).squeeze(-1)
    val_normal = (g3 * weights[:, :, None] * inside[:, :, None]).sum(dim=1)          # validate(), :380-382
    cam_d = torch.matmul(rot[None], rays_d[:, :, None]).squeeze(-1)
    render_depth = (mid_z * weights).sum(dim=1) * cam_d[:, 2]        # Q10

    gerr = (torch.linalg.norm(g3, ord=2, dim=-1) - 1.0) ** 2
    gradient_error = (relax_inside * gerr).sum() / (relax_inside.sum() + 1e-5)

    # first SDF zero-crossing (Q12, implicit_surface.py:181-216)
    sd = sdf.reshape(B, S)
    both = ((vmask[:, :-1] * vmask[:, 1:]) > 0).to(dt)
    cross = (sd[:, :-1] * sd[:, 1:] <= 0).to(dt)
    rank = torch.arange(S - 1, 0, -1, dtype=dt)                     # S-1 ... 1: earliest wins
    score = cross * rank[None, :] * both
    i0 = torch.argmax(score, dim=1, keepdim=True) if prev_idx is None else prev_idx.reshape(B, 1).long()
    i1 = i0 + 1
    mid_inside = (0.5 * (torch.gather(inside, 1, i0) + torch.gather(inside, 1, i1)) > 0.5).to(dt)
    mid_inside = mid_inside * (score.sum(dim=1, keepdim=True) > 0).to(dt)
    ga = torch.gather(g3, 1, i0[:, :, None].expand(-1, -1, 3))
    gb = torch.gather(g3, 1, i1[:, :, None].expand(-1, -1, 3))
    cosab = (ga * gb).sum(-1) / (torch.linalg.norm(ga, ord=2, dim=-1) * torch.linalg.norm(gb, ord=2, dim=-1) + 1e-8)
    mid_inside = mid_inside * (cosab > 0.5)
    s1, s2 = torch.gather(sd, 1, i0), torch.gather(sd, 1, i1)
    za, zb = torch.gather(mid_z, 1, i0), torch.gather(mid_z, 1, i1)
    z_cross = (s1 * zb - s2 * za) / (s1 - s2 + 1e-10)
    sdf_depth = z_cross * cam_d[:, None, 2] * mid_inside
    return {"alpha": alpha, "weights": weights, "weight_sum": weight_sum,
            "weight_max": torch.max(weights, dim=-1, keepdim=True)[0], "color_fine": color_fine, "normal": normal,
            "val_normal": val_normal, "render_depth": render_depth, "gradient_error": gradient_error,
            "inside_sphere": inside, "mid_inside_sphere": mid_inside, "sdf_depth": sdf_depth, "prev_idx": i0,
            "crossing_cos": cosab, "crossing_score": score.sum(dim=1, keepdim=True), "z_cross": z_cross}


# ----------------------------------------------------------------------------------------------
# checker: tolerance of the composited outputs (used by tests/ and by bench.py's parity block)
# ----------------------------------------------------------------------------------------------
def gradient_position_envelope(net, ref, mid_gpu, rays_o, rays_d, volumes, indexes):
    """Conditioning of d sdf / d x with respect to the SAMPLE POSITION.

    The sparse feature volumes are trilinear: their spatial derivative is piecewise constant along each axis and
    changes linearly with the other two coordinates at a rate of (feature difference) / voxel^2.  At the finest level
    of the benchmarked scene (704^3) one ulp of a sample depth (2.4e-7) moves the grid coordinate by 8e-5 of a voxel
    and the reference's OWN gradient by up to ~1e-4 of its scale (measured: tests/test_oracle_golden.py) — and the two
    paths' sample depths do differ by an ulp or two (fused vs separate multiply-add, stage windows).  The per-point
    gradient is therefore held to

        |g_gpu - g_ref|  <=  1e-4 * scale  +  2 * max_shift |g_ref(p + shift) - g_ref(p)|

    with the shifts +-(|mid_gpu - mid_ref| + 2.5e-7) along the ray and the same length along (1,1,1)/sqrt(3): the
    first-order image of the measured position difference plus one ulp of a grid coordinate.  Returns the (P,) envelope
    (max over components) for the evaluated samples (0 elsewhere)."""
    B, S = ref["mid_z_vals"].shape
    cm = ref["_compute_mask"]
    dz = (torch.as_tensor(mid_gpu).detach().cpu().double() - ref["mid_z_vals"].double()).abs().reshape(-1)[cm]
    dirs = rays_d[:, None, :].expand(B, S, 3).reshape(-1, 3)[cm].double()
    pts = ref["_pts"][cm]
    step = (dz * dirs.norm(dim=1) + 2.5e-7)[:, None]
    unit = dirs / dirs.norm(dim=1, keepdim=True)
    diag = torch.full_like(unit, 3.0 ** -0.5)
    g0 = ref["_grad"][cm].double()
    env = torch.zeros(pts.shape[0], dtype=torch.float64)
    for shift in (unit * step, -unit * step, diag * step):
        _, g = sdf_gradient(net, (pts.double() + shift).float(), volumes, indexes)
        env = torch.maximum(env, (g.double() - g0).abs().max(dim=1)[0])
    out = torch.zeros(B * S, dtype=torch.float64)
    out[cm] = env
    return out


def color_position_envelope(net, ref, mid_gpu, rays_o, rays_d, imgs, intrs, c2ws, features):
    """Conditioning of the blended colour with respect to the sample position / its projection.

    A sample projects to pixel coordinates of magnitude ~10^2..10^3, whose fp32 spacing is 3e-5..6e-5 px; the two paths
    form the projection in a different operation order (torch: inverse(c2w) @ p, K @ cam; the kernel: precomputed w2c
    rows) and their sample depths differ by an ulp or two, so the bilinear taps are taken ~1e-4 px apart — and the
    sampled images / feature maps change by O(1) per pixel (white noise in the synthetic scenes).  The per-point colour
    is therefore held to 1e-4 + 2 x (pooling-weight envelope + this envelope): the change of the ORACLE's colour when
    the point moves by +-(|mid_gpu - mid_ref| * |d| + 2.5e-7) along the ray and by the same length along (1,1,1)/sqrt 3
    (2.5e-7 in world units = 1.8e-4 px at the benchmarked geometry = 3 ulp of a pixel coordinate).
    Returns the (P,) envelope (max over RGB) for the evaluated samples (0 elsewhere)."""
    B, S = ref["mid_z_vals"].shape
    cm = ref["_compute_mask"]
    dz = (torch.as_tensor(mid_gpu).detach().cpu().double() - ref["mid_z_vals"].double()).abs().reshape(-1)[cm]
    dirs = rays_d[:, None, :].expand(B, S, 3).reshape(-1, 3)[cm].double()
    pts = ref["_pts"][cm]
    step = (dz * dirs.norm(dim=1) + 2.5e-7)[:, None]
    unit = dirs / dirs.norm(dim=1, keepdim=True)
    diag = torch.full_like(unit, 3.0 ** -0.5)
    fv, rd, mk = lookup_feature(pts, imgs, intrs, c2ws, features)
    c0 = blend(net, fv, rd, mk).double()
    env = torch.zeros(pts.shape[0], dtype=torch.float64)
    for shift in (unit * step, -unit * step, diag * step):
        fv1, rd1, mk1 = lookup_feature((pts.double() + shift).float(), imgs, intrs, c2ws, features)
        c1 = blend(net, fv1, rd1, mk1).double()
        env = torch.maximum(env, (c1 - c0).abs().max(dim=1)[0])
    out = torch.zeros(B * S, dtype=torch.float64)
    out[cm] = env
    return out


RAY_KEYS = ("color_fine", "render_depth", "sdf_depth", "normal", "val_normal", "weight_sum")


def composite_envelope(ref, sdf_g, grad_g, color_g, rays_o, rays_d, inv_s, rot0_inv, cos_anneal_ratio=1.0,
                       per_sample=False):
    """NeuS' alpha multiplies an SDF difference by inv_s (20 at init, 3000 in the sharpened goldens) before a
    sigmoid (implicit_surface.py:126-149), so a per-point SDF deviation of 3e-6 — thirty times inside the
    north-star's 1e-4 — moves a composited output by far more than 1e-4 of its scale.  The composited outputs
    are therefore held to

        |out_gpu - out_ref|  <=  1e-4 * scale  +  1.5 * sum_i |d out / d x_i| * |x_gpu_i - x_ref_i|

    where x runs over the per-point SDF, gradient and colour values (each separately asserted within 1e-4 of the
    reference), and the Jacobian is taken by float64 autograd through the ORACLE's compositing at the reference
    values.  In words: the compositing kernel itself adds at most 1e-4; the rest is the first-order image of
    per-point deviations that are already inside the tolerance.  ``ref`` = oracle render_core(return_stages=True).
    Returns ({key: (B,C) envelope}, {key: (B,C) oracle fp64 value})."""
    B, S = ref["mid_z_vals"].shape
    d = torch.float64
    x_ref = [ref["_sdf"].to(d), ref["_grad"].to(d), ref["_color"].reshape(-1, 3).to(d)]
    delta = [(torch.as_tensor(g).detach().cpu().to(d).reshape(x.shape) - x).abs()
             for g, x in zip((sdf_g, grad_g, color_g), x_ref)]
    xs = [x.clone().requires_grad_(True) for x in x_ref]
    with torch.enable_grad():
        comp = composite(rays_o.to(d), rays_d.to(d), ref["mid_z_vals"].to(d), ref["_dists"].to(d),
                           ref["_voxel_mask"].reshape(B, S).to(d), xs[0], xs[1], xs[2],
                           torch.as_tensor(inv_s).to(d), rot0_inv.to(d), cos_anneal_ratio, prev_idx=ref["_prev_idx"])
        env, val = {}, {}
        keys = list(RAY_KEYS) + (["alpha"] if per_sample else [])
        for k in keys:
            o = comp[k].reshape(B, -1)
            val[k] = o.detach()
            if k == "alpha":        # alpha_ij depends on sample (i,j) only: one pass gives every derivative
                gs = torch.autograd.grad(o.sum(), xs, retain_graph=True, allow_unused=True)
                e = sum(((g.abs() * dl).reshape(B, S, -1).sum(dim=2)) for g, dl in zip(gs, delta) if g is not None)
                env[k] = e
                continue
            cols = []
            for c in range(o.shape[1]):
                gs = torch.autograd.grad(o[:, c].sum(), xs, retain_graph=True, allow_unused=True)
                e = sum(((g.abs() * dl).reshape(B, -1).sum(dim=1)) for g, dl in zip(gs, delta) if g is not None)
                cols.append(e)
            env[k] = torch.stack(cols, dim=1)
        if per_sample:
            w = comp["weights"]
            cols = []
            for j in range(S):
                gs = torch.autograd.grad(w[:, j].sum(), xs, retain_graph=True, allow_unused=True)
                cols.append(sum(((g.abs() * dl).reshape(B, -1).sum(dim=1)) for g, dl in zip(gs, delta) if g is not None))
            env["weights"] = torch.stack(cols, dim=1)
            val["weights"] = w.detach()
    return env, val



# ----------------------------------------------------------------------------------------------
# MatchingField  (matching_field.py:18-141): the probe of the matching volume run for every pixel of every view
# ----------------------------------------------------------------------------------------------
def matching_depth_render(rays_o, rays_d, near, far, c2w, matching_volume, n_samples, t_rand=None):
    """matching_field.py:18-72.  near / far (B, n_windows); per window n_samples uniform depths (+ jitter
    (t_rand[:, i] - 0.5) * (far - near) / n when t_rand (B, n_windows) of raw U[0,1) draws is given), all windows
    sorted together, softmax of the trilinear probes -> expected depth.  Returns (render_depth (B,), occ_reg scalar,
    z_vals (B,S), density (B,S))."""
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    B = rays_o.shape[0]
    zs = []
    for i in range(near.shape[-1]):
        ns, fs = near[..., [i]], far[..., [i]]
        z = ns + (fs - ns) * torch.linspace(0.0, 1.0, n_samples)[None, :]
        if t_rand is not None:
            z = z + (t_rand[:, i:i + 1] - 0.5) * (fs - ns) / n_samples
        zs.append(z)
    z_vals, _ = torch.sort(torch.cat(zs, dim=-1), dim=-1)
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]).reshape(-1, 3)
    outside = (torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(B, -1) > 1.0).float()
    density = lookup_dense(pts, matching_volume, "bilinear").reshape(B, -1)
    weights = F.softmax(density, dim=-1)
    cam_d = torch.matmul(torch.inverse(c2w[None, :3, :3]), rays_d[:, :, None]).squeeze()
    render_depth = (z_vals * weights).sum(dim=1) * cam_d[:, 2]
    occ_reg = density[:, :6].mean() + (density * outside).sum() / (outside.sum() + 1e-10)
    return render_depth, occ_reg, z_vals, density


def matching_rays(intrs, c2ws, view, img_h, img_w, level):
    """matching_field.py:79-100: the pixel grid of a depth map at 1/level resolution and its rays in view `view`."""
    h, w = img_h // level, img_w // level
    tx = torch.linspace(0, img_w - 1, w)
    ty = torch.linspace(0, img_h - 1, h)
    py, px = torch.meshgrid(ty, tx, indexing="ij")
    px, py = px.reshape(-1), py.reshape(-1)
    pix = torch.stack([px, py, torch.ones_like(px)], dim=-1).float()
    cam = torch.matmul(intrs.inverse()[view, None, :3, :3], pix[:, :, None]).squeeze()
    d = cam / torch.linalg.norm(cam, ord=2, dim=-1, keepdim=True)
    d = torch.matmul(c2ws[view, None, :3, :3], d[:, :, None]).squeeze()
    o = c2ws[view, None, :3, 3].expand(d.shape)
    return o, d, px, py, h, w


def matching_windows(near_ori, far_ori, pre_z, ratio):
    """matching_field.py:107-121: a window of width (far - near) * ratio around pre_z, shifted into [near, far]."""
    sr = (far_ori - near_ori).squeeze() * ratio
    near = (pre_z - sr / 2).unsqueeze(1)
    far = (pre_z + sr / 2).unsqueeze(1)
    near = torch.where(far > far_ori, near - (far - far_ori), near)
    far = torch.where(near < near_ori, far + (near_ori - near), far)
    near = torch.clamp(near, near_ori.squeeze(), far_ori.squeeze())
    far = torch.clamp(far, near_ori.squeeze(), far_ori.squeeze())
    return near, far


def matching_field_forward(n_samples_depths, depth_res_levels, ipts, matching_volume, stage_idx, range_ratios,
                           pre_depths=None, perturb=False):
    """MatchingField.forward (matching_field.py:74-141): a depth map per view at the stage's resolution, bilinearly
    up-sampled to the image size, + the occupancy regulariser per view."""
    near_fars, c2ws, intrs = ipts["near_fars"], ipts["c2ws"], ipts["intrs"]
    src_idx = ipts["src_idx"] if "src_idx" in ipts else 0
    img_h, img_w = ipts["imgs"].shape[-2:]
    level = depth_res_levels[stage_idx]
    n = n_samples_depths[stage_idx]
    depths, occs = [], []
    for i in range(intrs.shape[0]):
        o, d, px, py, h, w = matching_rays(intrs, c2ws, i, img_h, img_w, level)
        near_ori, far_ori = near_fars[i].reshape(1, 2).split(split_size=1, dim=1)
        if pre_depths is not None:
            pre = pre_depths[i].detach()[(py.long(), px.long())]
            cam_d = torch.matmul(torch.inverse(c2ws[i, None, :3, :3]), d[:, :, None]).squeeze()
            pre_z = pre.reshape(-1) / cam_d[:, 2]
            n1, f1 = matching_windows(near_ori, far_ori, pre_z, range_ratios[stage_idx])
            n0, f0 = matching_windows(near_ori, far_ori, pre_z, range_ratios[stage_idx - 1])
            near, far = torch.cat([n1, n0], dim=1), torch.cat([f1, f0], dim=1)
        else:
            near, far = near_ori.repeat(o.shape[0], 1), far_ori.repeat(o.shape[0], 1)
        t_rand = None
        if perturb and (i == 0 or i == src_idx):
            t_rand = torch.cat([torch.rand([o.shape[0], 1]) for _ in range(near.shape[-1])], dim=1)
        rd, occ, _, _ = matching_depth_render(o, d, near, far, c2ws[i], matching_volume, n, t_rand)
        rd = F.interpolate(rd.reshape(1, 1, h, w), size=(img_h, img_w), mode="bilinear").squeeze(0).squeeze(0)
        depths.append(rd)
        occs.append(occ)
    return depths, occs


# ----------------------------------------------------------------------------------------------
# Volume  (volume.py:21-168): producers of the scene tensors
# ----------------------------------------------------------------------------------------------
def volume_voxel_size(dims, bounding=((-1.0, 1.0), (-1.0, 1.0), (-1.0, 1.0))):
    """volume.py:22-23: float64 numpy voxel size and origin of a (dx,dy,dz) grid spanning the bounding box."""
    import numpy as np
    b = np.array(bounding, dtype=np.float64)
    return (b[:, 1] - b[:, 0]) / (np.array(dims) - 1), b[:, 0]


def volume_init_coords(dims):
    """volume.py:21-33: integer grid coordinates (n,3) fp32, x-major."""
    g = torch.stack(torch.meshgrid(*[torch.arange(0, int(d)) for d in dims], indexing="ij")).float()
    return g.view(3, -1).permute(1, 0).contiguous()


def volume_up_sample(pre_coords, pre_feat, num=8):
    """volume.py:35-52: every voxel -> its 8 children at twice the resolution (coords * 2 + offset), features repeated.
    (The reference doubles ``pre_coords`` in place; this restatement does not mutate its input.)"""
    c2 = pre_coords * 2
    offs = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]],
                        dtype=c2.dtype)[:num]
    up_coords = (c2[:, None, :] + offs[None, :, :]).reshape(-1, 3)
    up_feat = pre_feat[:, None, :].expand(-1, num, -1).reshape(-1, pre_feat.shape[1])
    return up_coords, up_feat


def _volume_project(coords, voxel_size, origin, intrs, c2ws, h, w):
    world = coords * torch.tensor(voxel_size).type_as(coords)[None] + torch.tensor(origin).type_as(coords)[None]
    world = world[None].permute(0, 2, 1).contiguous()
    world = torch.cat([world, torch.ones_like(world[:, :1])], dim=1)                  # (1,4,n)
    cam = torch.matmul(torch.inverse(c2ws), world)
    img = torch.matmul(intrs, cam)[:, :3]
    xy = img[:, :2] / img[:, 2:]
    nx = xy[:, 0] / ((w - 1) / 2) - 1
    ny = xy[:, 1] / ((h - 1) / 2) - 1
    mask = (nx.abs() <= 1) & (ny.abs() <= 1) & (img[:, 2] > 0)                        # (nv,n)
    return torch.stack([nx, ny], dim=-1), mask, img[:, 2:]


def volume_back_proj(agg, feats, coords, voxel_size, origin, intrs, c2ws, stage_idx):
    """volume.py:54-97.  feats: coarse->fine list of (nv,c,h,w); agg = dict of agg_mlp weights ('0.weight', '0.bias',
    '2.weight', '2.bias').  Projects every voxel into every view, sums the bilinear samples of the feature scales
    stage_idx.., scores each view with the 4->8->1 MLP, masked softmax over the views -> (mean | 'variance') (n,2c),
    and the >= 2-view frustum mask."""
    nv, c, h, w = feats[-1].shape
    grid, mask, _ = _volume_project(coords, voxel_size, origin, intrs, c2ws, h, w)
    warp = 0
    for f in feats[stage_idx:]:
        warp = warp + F.grid_sample(f, grid.unsqueeze(1), padding_mode="zeros", align_corners=True).reshape(nv, c, -1)
    warp = warp.permute(0, 2, 1).contiguous()                                          # (nv,n,c)
    x = F.linear(F.elu(F.linear(warp, agg["0.weight"], agg["0.bias"])), agg["2.weight"], agg["2.bias"])
    x = x.masked_fill(mask.unsqueeze(-1) == 0, -1e9)
    wv = F.softmax(x, dim=0)
    mean = (warp * wv).sum(dim=0)
    var = ((warp * wv) ** 2).sum(dim=0) - (warp * wv).sum(dim=0) ** 2
    return torch.cat([mean, var], dim=1), mask.sum(dim=0) > 1


def volume_depth_filter_mask(depths, coords, voxel_size, origin, intrs, c2ws, depth_range):
    """volume.py:134-168: a voxel survives when its depth agrees (within depth_range) with the rendered depth map in at
    least two views.  depths: list of (h,w).  Returns the bool (n,) mask."""
    d = torch.stack(depths, dim=0).unsqueeze(1)
    nv, _, h, w = d.shape
    grid, mask, cz = _volume_project(coords, voxel_size, origin, intrs, c2ws, h, w)
    wd = F.grid_sample(d, grid.unsqueeze(1), padding_mode="zeros", align_corners=True).reshape(nv, 1, -1)
    valid = ((wd - cz).abs() < depth_range) & mask.unsqueeze(1)
    return (valid.sum(0) > 1).squeeze(0)


def volume_sparse2dense(feats, coords, dims, pre_volume=None):
    """volume.py:99-121: scatter (n,c) voxel values into a dense (1,c,D,H,W) volume; channel 0 of the empty voxels is
    the trilinearly 2x up-sampled previous volume.  Also the fp32 0/1 mask volume."""
    c = feats.shape[1]
    dense = torch.zeros([1, int(dims[0]), int(dims[1]), int(dims[2]), c])
    if pre_volume is not None:
        dense[..., :1] = F.interpolate(pre_volume, scale_factor=2, mode="trilinear").permute(0, 2, 3, 4, 1)
    maskv = torch.zeros([1, int(dims[0]), int(dims[1]), int(dims[2]), 1])
    loc = coords.to(torch.int64)
    dense[:, loc[:, 0], loc[:, 1], loc[:, 2]] = feats
    maskv[:, loc[:, 0], loc[:, 1], loc[:, 2]] = 1.0
    return dense.permute(0, 4, 1, 2, 3), maskv.permute(0, 4, 1, 2, 3)


def volume_get_index(coords, dims):
    """volume.py:123-132: int64 table, -1 = empty, entry = row of the voxel."""
    t = torch.full([int(dims[0]), int(dims[1]), int(dims[2])], -1, dtype=torch.int64)
    loc = coords.to(torch.int64)
    t[loc[:, 0], loc[:, 1], loc[:, 2]] = torch.arange(coords.shape[0], dtype=torch.int64)
    return t


# ----------------------------------------------------------------------------------------------
# training extras  (implicit_surface.py:172, 218-245; projector.py:560-645)
# ----------------------------------------------------------------------------------------------
def sdf_gradient_smooth(net: OracleNet, pts: torch.Tensor, volumes, indexes):
    """sdf_network.gradient (sdf_network.py:129-152): first-order gradient and `smooth` = d/dx of sum_j (d sdf / d x_j)
    = Hessian . (1,1,1), by double autograd like the reference.  Returns (grad (n,3), smooth (n,3)), detached."""
    with torch.enable_grad():
        x = pts.detach().clone().requires_grad_(True)
        y = sdf_only(net, x, volumes, indexes)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True)
        (sm,) = torch.autograd.grad(g, x, torch.ones_like(g))
    return g.detach(), sm.detach()


def warp_feature_maps(features):
    """implicit_surface.py:229-241: the 12-channel maps the patch loss compares = level-0 features, level-1 and
    level-2 features bilinearly up-sampled to the level-0 size (F.interpolate, align_corners=False).  (nv,12,H,W)."""
    f0 = features[0]
    f1 = F.interpolate(features[1], size=f0.shape[-2:], mode="bilinear")
    f2 = F.interpolate(features[2], size=f0.shape[-2:], mode="bilinear")
    return torch.cat([f0, f1, f2], dim=1)


def patch_homography(H, uv):
    """projector.py:631-645.  H (B,V,3,3), uv (B,121,2) -> (V, B*121, 2) warped pixel coordinates."""
    B, V = H.shape[0], H.shape[1]
    hom = torch.cat([uv, torch.ones_like(uv[..., :1])], dim=-1)              # (B,121,3)
    t = torch.einsum("bvik,bok->vboi", H, hom)                                # (V,B,121,3)
    t = t.reshape(V, -1, 3)
    return t[..., :2] / (t[..., 2:] + 1e-8)


def surface_patch_warp2(pts_sdf0, normals, images, intrs, c2ws, patch_size=11, pixel_shift=None):
    """projector.py:560-628.  pts_sdf0 (B,1,3) world points, normals (B,1,3) unit normals in the REFERENCE camera
    frame, images (nv,C,H,W).  Plane-induced homography of an 11x11 pixel patch around the projection of each point
    into every source view; -> ref values (1,B,121,C), source values (V,B,121,C)."""
    B = pts_sdf0.shape[0]
    K_inv = torch.inverse(intrs)
    R0t = c2ws[0, :3, :3].permute(1, 0).contiguous()
    xyz = torch.matmul(R0t, pts_sdf0.permute(0, 2, 1).contiguous())           # (B,3,1)
    xyz = xyz + (-torch.matmul(R0t, c2ws[0, :3, 3, None]))
    pts_ref = xyz
    proj = torch.matmul(intrs[0, :3, :3], xyz)
    disp = torch.matmul(normals, pts_ref)                                       # (B,1,1)
    K_ref_inv = K_inv[0, :3, :3]
    K_src = intrs[1:, :3, :3]
    V = K_src.shape[0]
    R_src = c2ws[1:, :3, :3].permute(0, 2, 1).contiguous()
    R_rel = torch.matmul(R_src, c2ws[0, :3, :3])
    C_rel = c2ws[0, :3, 3][None, ...] - c2ws[1:, :3, 3]
    tmp = torch.matmul(R_src, C_rel[..., None])                                 # (V,3,1)
    tmp = torch.matmul(tmp[None, ...].expand(B, V, 3, 1), normals.expand(B, V, 3)[..., None].permute(0, 1, 3, 2))
    tmp = R_rel[None, ...].expand(B, V, 3, 3) + tmp / (disp[..., None] + 1e-10)
    tmp = torch.matmul(K_src[None, ...].expand(B, V, 3, 3), tmp)
    Hom = torch.matmul(tmp, K_ref_inv[None, None, ...])
    px = proj[:, 0, 0] / (proj[:, 2, 0] + 1e-8)
    py = proj[:, 1, 0] / (proj[:, 2, 0] + 1e-8)
    pixels = torch.stack([px, py], dim=-1).float()
    hp = patch_size // 2
    total = (2 * hp + 1) ** 2
    off = torch.arange(-hp, hp + 1)
    off = torch.stack(torch.meshgrid(off, off, indexing="ij")[::-1], dim=-1).view(1, -1, 2).type_as(pixels)
    patch = pixels.view(B, 1, 2) + off.float()                                  # (B,121,2)
    ref_img, src_imgs = images[0], images[1:]
    h, w = ref_img.shape[-2:]
    grid = patch_homography(Hom, patch)
    if pixel_shift is not None:       # checker only: sensitivity of the bilinear taps to the last bits of a pixel coordinate
        grid = grid + torch.as_tensor(pixel_shift, dtype=grid.dtype)
        patch = patch + torch.as_tensor(pixel_shift, dtype=patch.dtype)
    grid = torch.stack([2 * grid[:, :, 0] / (w - 1) - 1.0, 2 * grid[:, :, 1] / (h - 1) - 1.0], dim=-1)
    sampled = F.grid_sample(src_imgs, grid.view(V, -1, 1, 2), align_corners=True)
    sampled = sampled.view(V, -1, B, total).permute(0, 2, 3, 1).contiguous()
    pg = torch.stack([2 * patch[:, :, 0] / (w - 1) - 1.0, 2 * patch[:, :, 1] / (h - 1) - 1.0], dim=-1)
    refv = F.grid_sample(ref_img[None, ...], pg.view(1, -1, 1, 2), align_corners=True)
    refv = refv.view(1, -1, B, total).permute(0, 2, 3, 1).contiguous()
    return refv, sampled, {"disp": disp.reshape(B), "pixels": pixels, "grid": grid.view(V, B, total, 2)}


def render_extras(net: OracleNet, rays_o, rays_d, z_vals, z_cross, inside, gradients, volumes, indexes, features,
                  intrs, c2ws, smooth=True):
    """The training-only tail of render_core (implicit_surface.py:172, 218-245): `smooth_error`, the surface point
    `pts_sdf0` (zero-crossing depth clamped to [0, max z_vals of the CHUNK], Q15), its unit normal in the reference
    camera frame (third MLP pass), and the warped 11x11 feature patches."""
    B = rays_o.shape[0]
    z0 = torch.where(z_cross < 0, torch.zeros_like(z_cross), z_cross)
    z0 = torch.where(z0 > torch.max(z_vals), torch.zeros_like(z0), z0)
    pts_sdf0 = rays_o[:, None, :] + rays_d[:, None, :] * z0[..., :, None]       # (B,1,3)
    _, g0 = sdf_gradient(net, pts_sdf0.reshape(-1, 3), volumes, indexes)
    g0 = g0.reshape(B, 1, 3)
    n0 = torch.linalg.norm(g0, ord=2, dim=-1, keepdim=True)
    n0 = torch.where(n0 <= 0, torch.ones_like(n0) * 1e-8, n0)
    g0 = g0 / n0
    g0 = torch.matmul(c2ws[0, :3, :3].permute(1, 0).contiguous()[None, ...],
                      g0.permute(0, 2, 1).contiguous()).permute(0, 2, 1).contiguous()
    refv, sampled, dbg = surface_patch_warp2(pts_sdf0, g0, warp_feature_maps(features), intrs, c2ws)
    out = {"ref_gray_val": refv, "sampled_gray_val": sampled, "_pts_sdf0": pts_sdf0.reshape(B, 3),
           "_normal_sdf0": g0.reshape(B, 3), "_disp": dbg["disp"], "_patch_grid": dbg["grid"]}
    return out


# ----------------------------------------------------------------------------------------------
# render_core  (implicit_surface.py:64-266; quirks Q2, Q5-Q12)
# ----------------------------------------------------------------------------------------------
def render_core(net: OracleNet, rays_o, rays_d, z_vals, volumes, indexes, mask_volumes, features,
                imgs, intrs, c2ws, cos_anneal_ratio=1.0, pts_random: Optional[torch.Tensor] = None,
                return_stages=False, extras=False):
    B, S = z_vals.shape
    sample_dist = 2.0 / net.n_samples[0]
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    dists = torch.cat([dists, torch.full((B, 1), sample_dist, dtype=z_vals.dtype)], dim=-1)
    mid_z = z_vals + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, S, 3).reshape(-1, 3)

    vmask = point_mask(pts, mask_volumes)                     # bool (P,)
    compute = vmask.clone()
    if int(compute.sum()) < 1:                                # Q6 (implicit_surface.py:88-89)
        compute[:10] = True
    vmask_f = vmask.to(torch.float32)

    P = pts.shape[0]
    sdf = torch.full((P, 1), 100.0)                           # Q7
    grad = torch.zeros(P, 3)
    color = torch.zeros(P, 3)
    view_mask = torch.zeros(P, imgs.shape[0] - 1, dtype=torch.bool)
    pv = pts[compute]
    s_v, g_v = sdf_gradient(net, pv, volumes, indexes)
    sdf[compute] = s_v
    grad[compute] = g_v
    fv, rd, mv = lookup_feature(pv, imgs, intrs, c2ws, features)
    color[compute] = blend(net, fv, rd, mv)
    view_mask[compute] = mv
    valid_mask = ((view_mask.reshape(B, S, -1).float().sum(dim=2) > 1).float().sum(dim=1, keepdim=True) > 8)  # Q11

    inv_s = torch.exp(net.variance * 10.0).clip(1e-6, 1e6)    # Q9
    comp = composite(rays_o, rays_d, mid_z, dists, vmask_f.reshape(B, S), sdf, grad, color, inv_s,
                     torch.inverse(c2ws[0, :3, :3]), cos_anneal_ratio)
    alpha, weights, weight_sum = comp["alpha"], comp["weights"], comp["weight_sum"]
    color_fine, normal, render_depth = comp["color_fine"], comp["normal"], comp["render_depth"]
    gradient_error, inside, mid_inside = comp["gradient_error"], comp["inside_sphere"], comp["mid_inside_sphere"]
    sdf_depth, i0, g3 = comp["sdf_depth"], comp["prev_idx"], grad.reshape(B, S, 3)

    # 1024 uniform random points -> sparse_sdf (implicit_surface.py:174-178); draws from the
    # global generator unless given, to keep the RNG stream of chunked validation in step (Q1)
    if pts_random is None:
        pts_random = torch.rand([1024, 3]) * 2 - 1
    rmask = point_mask(pts_random, mask_volumes)
    sdf_random = torch.zeros(pts_random.shape[0], 1)
    if bool(rmask.any()):
        sdf_random[rmask] = sdf_only(net, pts_random[rmask], volumes, indexes)

    out = {
        "color_fine": color_fine,
        "render_depth": render_depth,
        "sdf_depth": sdf_depth,
        "normal": normal,
        "valid_mask": valid_mask,
        "sparse_sdf": torch.cat([sdf_random, sdf]),
        "mid_z_vals": mid_z,
        "gradients": g3,
        "s_val": (1.0 / inv_s).reshape(1, 1).expand(P, 1),
        "weights": weights,
        "weight_sum": weight_sum,
        "weight_max": comp["weight_max"],
        "gradient_error": gradient_error,
        "inside_sphere": inside,
        "mid_inside_sphere": mid_inside,
    }
    if extras:
        # smooth_error (:172): masked-out samples carry smooth = 0 (:99,:108)
        smooth = torch.zeros(P, 3)
        smooth[compute] = sdf_gradient_smooth(net, pv, volumes, indexes)[1]
        out["smooth_error"] = (torch.linalg.norm(smooth, ord=2, dim=-1).reshape(B, S) * inside).sum() / (inside.sum() + 1e-5)
        out.update(render_extras(net, rays_o, rays_d, z_vals, comp["z_cross"], inside, g3, volumes, indexes, features,
                                 intrs, c2ws))
    if return_stages:
        out.update({"_pts": pts, "_voxel_mask": vmask, "_compute_mask": compute, "_sdf": sdf,
                    "_color": color.reshape(B, S, 3), "_view_mask": view_mask, "_alpha": alpha,
                    "_prev_idx": i0, "_dists": dists, "_grad": grad, "_val_normal": comp["val_normal"],
                    "_crossing_cos": comp["crossing_cos"], "_crossing_score": comp["crossing_score"]})
    return out


def render(net: OracleNet, rays_o, rays_d, near, far, matching_volume, volumes, indexes, mask_volumes,
           imgs, features, intrs, c2ws, cos_anneal_ratio=1.0, t_rand=None, pts_random=None,
           return_stages=False, extras=False):
    """ImplicitSurface.render (implicit_surface.py:268-335).  With t_rand/pts_random = None the
    jitter and the random points are drawn from torch's global CPU generator in the reference's
    order (Q1), so ``torch.manual_seed(s); render(...)`` matches the reference stream."""
    if near.shape[0] == 1 and rays_o.shape[0] != 1:
        near = near.expand(rays_o.shape[0], 1)
        far = far.expand(rays_o.shape[0], 1)
    if t_rand is None and net.perturb > 0:
        t_rand = draw_t_rand(rays_o.shape[0], len(net.n_samples))
    z, surf_z = sample_z(net, rays_o, rays_d, near, far, matching_volume, t_rand)
    out = render_core(net, rays_o, rays_d, z, volumes, indexes, mask_volumes, features, imgs, intrs, c2ws,
                      cos_anneal_ratio, pts_random, return_stages, extras)
    if return_stages:
        out["_z_vals"] = z
        out["_surf_z"] = surf_z
    return out


def validate_image(net: OracleNet, rays_o, rays_d, near, far, matching_volume, volumes, indexes,
                   mask_volumes, imgs, features, intrs, c2ws, hw, cos_anneal_ratio=1.0, chunk=256):
    """The image part of ImplicitSurface.validate (implicit_surface.py:366-400): 256-ray chunks."""
    import numpy as np
    if near.shape[0] == 1:
        near = near.expand(rays_o.shape[0], 1)
        far = far.expand(rays_o.shape[0], 1)
    rgb, nrm, sd, rd = [], [], [], []
    for o, d, n, f in zip(rays_o.split(chunk), rays_d.split(chunk), near.split(chunk), far.split(chunk)):
        r = render(net, o, d, n, f, matching_volume, volumes, indexes, mask_volumes, imgs, features, intrs,
                   c2ws, cos_anneal_ratio)
        rgb.append(r["color_fine"])
        nn_ = (r["gradients"] * r["weights"][:, :, None] * r["inside_sphere"][..., None]).sum(dim=1)
        nrm.append(nn_.numpy())
        sd.append(r["sdf_depth"].numpy())
        rd.append(r["render_depth"].numpy())
    h, w = int(hw[0]), int(hw[1])
    color = torch.cat(rgb, dim=0)
    rot = np.linalg.inv(c2ws[0, :3, :3].numpy())
    normal = np.concatenate(nrm, axis=0)
    return {
        "color_fine": color,
        "img_fine": (color.numpy().reshape(h, w, 3) * 256).clip(0, 255),
        "normal_img": (np.matmul(rot[None], normal[:, :, None]).reshape(h, w, 3) * 128 + 128).clip(0, 255),
        "sdf_depth": np.concatenate(sd, axis=0).reshape(h, w),
        "render_depth": np.concatenate(rd, axis=0).reshape(h, w),
    }


def sdf_grid(net: OracleNet, volumes, indexes, bound_min, bound_max, resolution,
             x_range=None, y_range=None, z_range=None, block=64):
    """u = -sdf on the linspace grid of extract_geometry (implicit_surface.py:337-351), dense (Q16).
    Optional index ranges restrict to a sub-box; returns u of that sub-box."""
    X = torch.linspace(float(bound_min[0]), float(bound_max[0]), resolution)
    Y = torch.linspace(float(bound_min[1]), float(bound_max[1]), resolution)
    Z = torch.linspace(float(bound_min[2]), float(bound_max[2]), resolution)
    xr = x_range or (0, resolution)
    yr = y_range or (0, resolution)
    zr = z_range or (0, resolution)
    X, Y, Z = X[xr[0]:xr[1]], Y[yr[0]:yr[1]], Z[zr[0]:zr[1]]
    u = torch.zeros(len(X), len(Y), len(Z))
    with torch.no_grad():
        for xi in range(0, len(X), block):
            for yi in range(0, len(Y), block):
                for zi in range(0, len(Z), block):
                    xs, ys, zs = X[xi:xi + block], Y[yi:yi + block], Z[zi:zi + block]
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                    p = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
                    v = -sdf_only(net, p, volumes, indexes).reshape(len(xs), len(ys), len(zs))
                    u[xi:xi + len(xs), yi:yi + len(ys), zi:zi + len(zs)] = v
    return u
