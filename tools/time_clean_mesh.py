"""Timing of the GPU mesh cleaning on a validate()-sized problem: a 512^3 marching-cubes mesh against 3 masks of
576 x 800, ray grid upscaled x2 (the reference's defaults)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from surf_b200 import clean_mesh as CM, mesh, synthetic

DEV = "cuda:0"
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ax = torch.linspace(-1, 1, res, device=DEV)
x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
u = 0.55 - torch.sqrt(x ** 2 + (y * 1.2) ** 2 + z ** 2) + 0.03 * torch.sin(25 * x) * torch.sin(21 * y) * torch.sin(17 * z)
del x, y, z
v, t = mesh.marching_cubes_device(u.contiguous(), 0.0)
del u
v = v / (res - 1) * 2.0 - 1.0
H, W, nv = 576, 800, 3
intrs, c2ws, _, _ = synthetic.make_cameras(nv, H, W)
yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
masks = torch.stack([(((xx - W / 2 - 20 * i) / (0.3 * W)) ** 2 + ((yy - H / 2) / (0.4 * H)) ** 2) <= 1.0 for i in range(nv)]).float().to(DEV)
print("mesh: %d vertices, %d faces" % (v.shape[0], t.shape[0]))


def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


vf = v.float()
ms, dil = ev(lambda: CM.dilate_masks(masks > 0.5, 11)); print("dilate 3 x 576 x 800, disk(11): %.3f ms" % ms)
ms, cnt = ev(lambda: CM.vertex_visibility(vf, dil, intrs, c2ws)); print("vertex visibility: %.3f ms" % ms)
f1 = t.long()[(cnt > 1)[t.long()].all(-1)]
ms, (hit, missed) = ev(lambda: CM.first_hit_faces(vf, f1, masks, intrs, c2ws, 2)); print("first hits, %d faces, 3 x 1152 x 1600 rays: %.3f ms" % (f1.shape[0], ms))
f2 = f1[hit]
ms, (lab, keep) = ev(lambda: CM.face_components(f2, 500)); print("components, %d faces: %.3f ms" % (f2.shape[0], ms))
t0 = time.perf_counter()
ov, of, st = CM.clean_mesh(v, t, masks, intrs, c2ws, return_stages=True)
torch.cuda.synchronize()
print("clean_mesh end to end (incl. host copies): %.1f ms" % ((time.perf_counter() - t0) * 1e3), st)
