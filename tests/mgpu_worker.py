"""Worker of tests/test_gpu_multi.py, launched under torchrun with one rank per GPU (NCCL): the north-star multi-GPU
workload — ONE image ray-sharded over the ranks + all-gather, the SDF grid x-slab sharded + all-gather (+ the mesh) —
must equal the single-GPU result bit for bit (no cross-rank reduction exists)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                                  # noqa: E402
import torch.distributed as dist              # noqa: E402

from helpers import load_golden, scene_from_recipe      # noqa: E402
from surf_b200 import conf, dist as sdist, mesh, synthetic   # noqa: E402
from surf_b200.modules.implicit_surface import ImplicitSurface   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    try:
        g = load_golden("validate_24x32")
        sc = scene_from_recipe(g["recipe"])
        m = ImplicitSurface(conf.default_implicit_surface_conf())
        m.load_state_dict(g["sd"])
        m = m.to(dev)
        d = sc.to(dev)
        ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
        # the 48x64 image of that scene minus 100 rays (2 972 rays = 12 chunks, the last one ragged)
        rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
        rays_o, rays_d = rays_o[:rays_o.shape[0] - 100].to(dev), rays_d[:rays_d.shape[0] - 100].to(dev)
        n = rays_o.shape[0]
        near, far = d.near, d.far
        torch.manual_seed(77)
        single = m.render_image(ps, rays_o, rays_d, near, far)                 # jitter drawn from the host stream
        torch.manual_seed(77)
        sharded = sdist.render_image_sharded(m, ps, rays_o, rays_d, near, far)
        ok = True
        for k in sdist.RECORD_KEYS:
            same = torch.equal(sharded[k].reshape(single[k].shape), single[k])
            ok = ok and same
            if not same:
                print("rank %d: %s differs (max %.3e)" % (rank, k, float((sharded[k].reshape(single[k].shape) - single[k]).abs().max())))
        # persistent-buffer gather (the bench's per-image call)
        r0, r1 = sdist.shard_rays(n, rank, world)
        ig = sdist.ImageGather(n, dev)
        loc = {k: single[k][r0:r1] for k in sdist.RECORD_KEYS}
        for _ in range(2):
            got = ig(loc)
            ok = ok and all(torch.equal(got[k].reshape(single[k].shape), single[k]) for k in sdist.RECORD_KEYS)
        # SDF grid: x-slabs + all-gather == one-GPU grid; the mesh of the gathered grid == the one-GPU mesh
        res = 96
        u1 = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], res)
        x0, x1 = sdist.shard_planes(res, rank, world)
        ug = sdist.gather_grid(m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], res, x_range=(x0, x1)), res)
        ok = ok and torch.equal(ug, u1)
        v1, t1 = mesh.marching_cubes_device(u1, 0.0)
        vg, tg = mesh.marching_cubes_device(ug, 0.0)
        ok = ok and torch.equal(v1, vg) and torch.equal(t1, tg) and t1.shape[0] > 1000
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("MGPU_RESULT world=%d ok=%d rays=%d tris=%d" % (world, int(flag.item()), n, int(t1.shape[0])))
        ps.destroy()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
