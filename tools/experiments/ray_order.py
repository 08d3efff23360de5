"""Experiment: does a 2-D tiled ray order (instead of image rows) help the gather locality of the render kernels?
Rays are independent, so the image is the same up to the permutation (the empty-chunk fallback aside)."""
import sys
sys.path.insert(0, "/root/repo")
import torch
import bench
from surf_b200 import synthetic

dev = "cuda:0"
H, W = 576, 800
sc = synthetic.make_scene(3, H, W, 88, seed=1, device=dev)
m = bench.build_net(dev)
ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
o, d, hw = synthetic.image_rays(sc, 1)
near, far = sc.near, sc.far
n = o.shape[0]
torch.manual_seed(1)
t = m.draw_chunk_randoms(n).to(dev)


def timed(fn, k=3):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


idx = torch.arange(n, device=dev).reshape(H, W)
for th, tw in ((1, 0), (8, 32), (16, 16), (4, 50), (32, 8), (64, 50)):
    if tw == 0:
        perm = idx.reshape(-1)
        name = "rows"
    else:
        perm = idx.reshape(H // th, th, W // tw, tw).permute(0, 2, 1, 3).reshape(-1)
        name = "%dx%d tiles" % (th, tw)
    oo, dd, tt = o[perm].contiguous(), d[perm].contiguous(), t[perm].contiguous()
    ms = timed(lambda: m.render_image(ps, oo, dd, near, far, t_rand=tt))
    print("%-12s %.2f ms / image" % (name, ms))
