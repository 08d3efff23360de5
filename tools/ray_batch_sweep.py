import os, sys
sys.path.insert(0, "/root/repo")
import torch, bench
from surf_b200 import synthetic, _lib
sc = synthetic.make_scene(3, 576, 800, 88, seed=1, device="cuda")
m = bench.build_net("cuda")
ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
near, far = sc.near, sc.far
rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.matching_volume = [], [], [], None
torch.cuda.empty_cache()
torch.manual_seed(1)
t = m.draw_chunk_randoms(rays_o.shape[0]).cuda()
for rb in [int(a) for a in sys.argv[1:]] or (1 << 16, 1 << 17, 1 << 18, 460800, 1 << 15):
    m.ray_batch = rb
    for _ in range(2):
        m.render_image(ps, rays_o, rays_d, near, far, t_rand=t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = m.render_image(ps, rays_o, rays_d, near, far, t_rand=t)
    e1.record(); torch.cuda.synchronize()
    print("ray_batch %7d: %.2f ms / image, peak mem %.1f GB" % (rb, e0.elapsed_time(e1) / 3, torch.cuda.max_memory_allocated() / 2**30), flush=True)
