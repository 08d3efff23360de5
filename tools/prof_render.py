"""A short render of 16384 coherent rays of the bench scene (for ncu captures of the per-ray kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from surf_b200 import _lib, synthetic
import bench
sc = synthetic.make_scene(3, 576, 800, 88, seed=1, device="cuda")
m = bench.build_net("cuda")
m.mlp_mode = int(os.environ.get("MLP_MODE", "1"))
ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
n = int(os.environ.get('N_RAYS', '65536'))
sel = slice(200 * 800, 200 * 800 + n)
torch.manual_seed(0)
t_rand = m.draw_chunk_randoms(n)
for _ in range(int(os.environ.get('REPS', '2'))):
    out = m.render_image(ps, rays_o[sel], rays_d[sel], sc.near, sc.far, t_rand=t_rand)
if os.environ.get("EXTRAS", "0") == "1":      # the training extras + marching cubes on a small grid, for their captures
    from surf_b200 import mesh
    o2, d2 = rays_o[sel][:2048].contiguous(), rays_d[sel][:2048].contiguous()
    m.render(o2, d2, sc.near.expand(2048, 1), sc.far.expand(2048, 1), ps, None, None, None, None, None, None, sc.intrs,
             sc.c2ws, 1.0, None)
    u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 256)
    mesh.marching_cubes_device(u, 0.0)
torch.cuda.synchronize()
print("ok")
