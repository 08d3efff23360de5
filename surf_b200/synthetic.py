"""Synthetic DTU-shaped scene generator.

There is no dataset in the build container or on the GPU box, and the
upstream that produces the renderer's inputs (FPN + torchsparse cost-volume
regularisation, surf.py:80-131) is out of scope, so benches and parity tests
feed the hot path with synthetic tensors that have exactly the types, layouts
and occupancy pattern ``SuRF.build_volumes`` emits:

  volumes[l]        (nvox_l, 7)  fp32   surf.py:119
  sparse_idxes[l]   (N,N,N)      int64, -1 = empty, else row of volumes[l]   volume.py:123-132
  mask_volumes[l]   (1,1,N,N,N)  fp32 0/1                                    volume.py:112-119
  matching_volume   (1,1,N3,N3,N3) fp32 logits at the finest level           volume.py:105-110
  features[i]       (nv,4,H/2^i,W/2^i) fp32 NCHW                             feature_network.py:178
  imgs (nv,3,H,W) in [0,1), intrs/c2ws (nv,4,4), rays on the reference view  dtu.py:383-467

All lists are returned in *renderer order*, i.e. already reversed the way
``SuRF.forward`` hands them to ``ImplicitSurface`` (surf.py:159): volume lists
fine->coarse, feature lists high-res->low-res.

World = cube [-1,1]^3; the geometric-init SDF is a sphere of radius 0.5, so the
sparse shells are centred on that sphere and rays really hit a surface inside
the finest shell.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch

RANGE_RATIOS = (1.0, 0.4, 0.1, 0.01)      # confs/surf.conf: model.range_ratios
FOCAL_AT_800 = 1446.0                      # DTU 2892.33 px @1600 wide -> 1446 @800
CAM_DIST = 2.0


@dataclass
class Scene:
    nv: int
    H: int
    W: int
    base: int
    imgs: torch.Tensor
    intrs: torch.Tensor
    c2ws: torch.Tensor
    near: torch.Tensor            # (1,1)
    far: torch.Tensor             # (1,1)
    matching_volume: torch.Tensor
    volumes: List[torch.Tensor] = field(default_factory=list)        # fine -> coarse
    sparse_idxes: List[torch.Tensor] = field(default_factory=list)   # fine -> coarse
    mask_volumes: List[torch.Tensor] = field(default_factory=list)   # fine -> coarse
    features: List[torch.Tensor] = field(default_factory=list)       # high-res -> low-res

    @property
    def device(self):
        return self.imgs.device

    def to(self, device):
        mv = lambda t: t.to(device)
        return Scene(self.nv, self.H, self.W, self.base, mv(self.imgs), mv(self.intrs), mv(self.c2ws),
                     mv(self.near), mv(self.far), mv(self.matching_volume),
                     [mv(t) for t in self.volumes], [mv(t) for t in self.sparse_idxes],
                     [mv(t) for t in self.mask_volumes], [mv(t) for t in self.features])

    def render_args(self):
        """Positional scene arguments of ImplicitSurface.render after (rays_o, rays_d, near, far)."""
        return (self.matching_volume, self.volumes, self.sparse_idxes, self.mask_volumes,
                self.imgs, self.features, self.features, self.intrs, self.c2ws)

    def voxel_counts(self):
        return [int(v.shape[0]) for v in self.volumes]


def make_cameras(nv, H, W, device="cpu"):
    """Reference view at (0,0,-2) looking down +z, sources displaced by 0.25 in x (then y)."""
    f = FOCAL_AT_800 * (W / 800.0)
    K = torch.eye(4, dtype=torch.float32)
    K[0, 0] = f
    K[1, 1] = f
    K[0, 2] = W / 2.0
    K[1, 2] = H / 2.0
    intrs = K[None].repeat(nv, 1, 1)
    c2ws = torch.eye(4, dtype=torch.float32)[None].repeat(nv, 1, 1)
    c2ws[:, 2, 3] = -CAM_DIST
    offs = [(0.0, 0.0), (0.25, 0.0), (-0.25, 0.0), (0.0, 0.25), (0.0, -0.25),
            (0.25, 0.25), (-0.25, -0.25), (0.25, -0.25), (-0.25, 0.25)]
    for v in range(nv):
        c2ws[v, 0, 3] = offs[v % len(offs)][0]
        c2ws[v, 1, 3] = offs[v % len(offs)][1]
    near = torch.tensor([[0.95 * (CAM_DIST - 1.0)]], dtype=torch.float32)
    far = torch.tensor([[1.05 * (CAM_DIST + 1.0)]], dtype=torch.float32)
    return intrs.to(device), c2ws.to(device), near.to(device), far.to(device)


def _axis(n, device):
    # voxel centres are laid out align_corners=True style (volume.py:23,64)
    return torch.arange(n, dtype=torch.float32, device=device) * (2.0 / (n - 1)) - 1.0


def _frustum_mask(n, intrs, c2ws, H, W, device):
    """Voxel centre visible in >= 2 views (same test as volume.py:78,95)."""
    ax = _axis(n, device)
    w2c = torch.inverse(c2ws.cpu()).to(device)
    cnt = torch.zeros((n, n, n), dtype=torch.int32, device=device)
    X = ax[:, None, None]
    Y = ax[None, :, None]
    Z = ax[None, None, :]
    for v in range(intrs.shape[0]):
        R = w2c[v]
        cx = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Z + R[0, 3]
        cy = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Z + R[1, 3]
        cz = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Z + R[2, 3]
        K = intrs[v]
        u = (K[0, 0] * cx + K[0, 2] * cz) / cz
        w_ = (K[1, 1] * cy + K[1, 2] * cz) / cz
        nx = u / ((W - 1) / 2.0) - 1.0
        ny = w_ / ((H - 1) / 2.0) - 1.0
        cnt += ((nx.abs() <= 1) & (ny.abs() <= 1) & (cz > 0)).to(torch.int32)
    return cnt > 1


def _shell_mask(n, parent, thresh, device, slab=64):
    """Children of ``parent`` voxels (2x up-sample, volume.py:35-52) whose centre lies within
    ``thresh`` of the r=0.5 sphere (emulates depth_filtering, volume.py:162-163)."""
    ax = _axis(n, device)
    out = torch.empty((n, n, n), dtype=torch.bool, device=device)
    y2 = (ax * ax)[None, :, None]
    z2 = (ax * ax)[None, None, :]
    for x0 in range(0, n, slab):
        x1 = min(n, x0 + slab)
        x2 = (ax[x0:x1] ** 2)[:, None, None]
        r = torch.sqrt(x2 + y2 + z2)
        par = parent[x0 // 2:(x1 + 1) // 2]
        par = par.repeat_interleave(2, 0)[(x0 % 2):(x0 % 2) + (x1 - x0)]
        par = par.repeat_interleave(2, 1).repeat_interleave(2, 2)
        out[x0:x1] = par & ((r - 0.5).abs() < thresh)
    return out


def make_scene(nv=3, H=576, W=800, base=88, seed=1, device="cpu", n_levels=4, feat_ch=7,
               range_ratios=RANGE_RATIOS, level0="frustum") -> Scene:
    """scene(nv, H, W, base, seed) of SURVEY.md §8d.  Deterministic for a given (args, device type).
    ``level0``: "frustum" = the coarsest mask is "voxel centre visible in >= 2 views" of THIS camera rig, as §8d's
    formula says (0.16 M / 1.3 M / 4.9 M / 4.6 M voxels at base 88); "all" = every coarsest voxel occupied, what a real
    scene with cameras all around gives and what §8d's voxel counts (0.68 M / 5.3 M / 8.6 M / 6.4 M) correspond to."""
    device = torch.device(device)
    g = torch.Generator(device=device)
    intrs, c2ws, near, far = make_cameras(nv, H, W, device)
    base_range = float(far - near)

    masks, vols, idxs = [], [], []
    parent = None
    for l in range(n_levels):
        n = base * (2 ** l)
        if l == 0 and level0 == "all":
            m = torch.ones((n, n, n), dtype=torch.bool, device=device)
        elif l == 0:
            m = _frustum_mask(n, intrs, c2ws, H, W, device)
        else:
            m = _shell_mask(n, parent, base_range * range_ratios[l], device)
        parent = m
        flat = m.reshape(-1)
        # running voxel number, in slabs of 2^28 entries (1408^3 = 2.8e9 entries: int64 table of 22 GB, built in place;
        # single torch reductions / scans over more than 2^31 elements are not reliable)
        idx = torch.empty(flat.shape[0], dtype=torch.int64, device=device)
        nvox = 0
        for a in range(0, flat.shape[0], 1 << 28):
            f = flat[a:a + (1 << 28)]
            c = torch.cumsum(f, 0, dtype=torch.int64)
            c += nvox - 1
            c.masked_fill_(~f, -1)
            idx[a:a + (1 << 28)] = c
            nvox += int(f.sum())
            del c
        idx = idx.reshape(n, n, n)
        g.manual_seed(seed + 100 + l)
        vol = torch.randn((nvox, feat_ch), generator=g, device=device, dtype=torch.float32) * 0.1
        masks.append(m.to(torch.float32).reshape(1, 1, n, n, n))
        idxs.append(idx)
        vols.append(vol)

    # matching logits at the finest level: peaked on the r=0.5 sphere + noise
    n = base * (2 ** (n_levels - 1))
    ax = _axis(n, device)
    g.manual_seed(seed + 200)
    matching = torch.empty((n, n, n), dtype=torch.float32, device=device)
    y2 = (ax * ax)[None, :, None]
    z2 = (ax * ax)[None, None, :]
    slab = 64
    for x0 in range(0, n, slab):
        x1 = min(n, x0 + slab)
        r = torch.sqrt((ax[x0:x1] ** 2)[:, None, None] + y2 + z2)
        noise = torch.randn(r.shape, generator=g, device=device, dtype=torch.float32)
        matching[x0:x1] = -40.0 * (r - 0.5).abs() + 0.1 * noise
    matching = matching.reshape(1, 1, n, n, n)

    g.manual_seed(seed + 300)
    imgs = torch.rand((nv, 3, H, W), generator=g, device=device, dtype=torch.float32)
    feats = []
    for i in range(4):
        g.manual_seed(seed + 400 + i)
        feats.append(torch.randn((nv, 4, H // (2 ** i), W // (2 ** i)), generator=g, device=device,
                                 dtype=torch.float32))

    return Scene(nv, H, W, base, imgs, intrs, c2ws, near, far, matching,
                 vols[::-1], idxs[::-1], masks[::-1], feats)


def make_rays(scene: Scene, pixels_x, pixels_y):
    """Unit rays through pixels of the reference view (datasets/dtu.py:428-432)."""
    dev = scene.device
    px = torch.as_tensor(pixels_x, dtype=torch.float32, device=dev).reshape(-1)
    py = torch.as_tensor(pixels_y, dtype=torch.float32, device=dev).reshape(-1)
    p = torch.stack([px, py, torch.ones_like(py)], dim=-1)
    Kinv = torch.inverse(scene.intrs[0, :3, :3].cpu()).to(dev)
    p = torch.matmul(Kinv[None], p[:, :, None]).squeeze(-1)
    d = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    d = torch.matmul(scene.c2ws[0, None, :3, :3], d[:, :, None]).squeeze(-1)
    o = scene.c2ws[0, None, :3, 3].expand(d.shape).contiguous()
    return o, d.contiguous()


def image_rays(scene: Scene, res_level=1):
    """All rays of the (H//res_level, W//res_level) validation grid, row-major (dtu.py:416-422)."""
    h, w = scene.H // res_level, scene.W // res_level
    tx = torch.linspace(0, scene.W - 1, w)
    ty = torch.linspace(0, scene.H - 1, h)
    py, px = torch.meshgrid(ty, tx, indexing="ij")
    o, d = make_rays(scene, px.reshape(-1), py.reshape(-1))
    return o, d, (h, w)


def random_pixel_rays(scene: Scene, n_rays, seed=2):
    g = torch.Generator().manual_seed(seed)
    px = torch.randint(0, scene.W, (n_rays,), generator=g)
    py = torch.randint(0, scene.H, (n_rays,), generator=g)
    return make_rays(scene, px, py)
