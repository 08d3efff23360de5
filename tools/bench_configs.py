#!/usr/bin/env python
"""Secondary benchmark configurations of BASELINE.json (configs[3], configs[4]); bench.py measures the headline
(configs[1]) and the 512^3 grid (configs[2]).  Prints one JSON line per configuration (rank 0).

    python tools/bench_configs.py cfgD [--steps K]          # training-shape render, all 18 training keys
    python tools/bench_configs.py cfgE [--steps K]          # Tanks-and-Temples-shaped image, reduced-precision MLP mode
    torchrun --nproc-per-node N ... tools/bench_configs.py cfgD|cfgE

cfgD: 4 x scene(5, 480, 640, base 64, seed 10+s), 512 random-pixel rays each, train-mode render()
(all 18 keys of the reference incl. the third MLP pass, the 11x11 patch warps and the second-order smooth_error); N ranks: rank r <-> scene r // 2,
half r % 2 of its rays (scene-major, no collective).
cfgE: scene(5, 1080, 1920, base 176 -> 1408^3 finest level, seed 20), n_samples = [64,32,16,16] (128 / ray), one image
of 2 073 600 rays in the opt-in one-MMA mode (tolerance 1e-2), ray-sharded over the ranks + NCCL all-gather.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from surf_b200 import _lib, conf, dist as sdist, synthetic  # noqa: E402


def timed(fn, steps, world, dev):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cfgD", "cfgE"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--base", type=int, default=0, help="override the coarsest volume dim (memory-limited boxes)")
    args = ap.parse_args()
    rank, world, local = bench.dist_env()
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    line = None
    if args.config == "cfgD":
        base = args.base or 64
        m = bench.build_net(dev)
        m.train()
        n_sc = 4
        per_scene = 512
        # rank r <-> scene r // 2, half r % 2 (world 8); fewer ranks take several scenes
        jobs = []                       # (scene id, first ray, last ray)
        units = [(s, h) for s in range(n_sc) for h in range(2)]
        for u_i, (s, h) in enumerate(units):
            if u_i % world == rank:
                jobs.append((s, h * per_scene // 2, (h + 1) * per_scene // 2))
        # a rank that holds both halves of a scene renders it in one call (the calls are launch-bound)
        merged = []
        for s, a, b in jobs:
            if merged and merged[-1][0] == s and merged[-1][2] == a:
                merged[-1] = (s, merged[-1][1], b)
            else:
                merged.append((s, a, b))
        jobs = merged
        scenes = {}
        for s, a, b in jobs:
            if s not in scenes:
                sc = synthetic.make_scene(5, 480, 640, base, seed=10 + s, device=dev)
                ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features,
                               sc.intrs, sc.c2ws)
                o, d = synthetic.random_pixel_rays(sc.to("cpu") if False else sc, per_scene, seed=30 + s)
                scenes[s] = (ps, o.to(dev), d.to(dev), sc.near.expand(per_scene, 1).contiguous(),
                             sc.far.expand(per_scene, 1).contiguous(), sc.intrs, sc.c2ws)
                sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.matching_volume = [], [], [], None
        torch.manual_seed(7)
        keys = None

        def step():
            nonlocal keys
            for s, a, b in jobs:
                ps, o, d, near, far, K, c2w = scenes[s]
                out = m.render(o[a:b], d[a:b], near[a:b], far[a:b], ps, None, None, None, None, None, None, K, c2w, 1.0, None)
                keys = sorted(k for k in out if not k.startswith("_"))
        for _ in range(max(3, args.warmup)):
            step()
        l0 = _lib.launch_count()
        ms = timed(step, args.steps, world, dev)
        launches = _lib.launch_count() - l0
        total = n_sc * per_scene
        line = {"metric": "rays_per_sec", "value": total * args.steps / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "dtype": "f32 (tcgen05 fp16 hi/lo split, fp32 accumulate)", "data": "synthetic",
                "config": {"workload": "cfgD: training-shape render 480x640, 5 views (4 src), 4 scenes x 512 rays per "
                                       "batch, volumes %d->%d, train-mode render() incl. patch warps" % (base, base * 8),
                           "parallelism": "scene-major: rank r <-> scene r // 2, half r % 2" if world > 1 else "1 GPU"},
                "gpu_launches": int(launches), "output_keys": keys,
                "missing_vs_reference": ["autograd through the render (outputs are detached)"]}
    else:
        base = args.base or 176
        c = conf.default_implicit_surface_conf()
        c.put("render.n_samples", [64, 32, 16, 16])
        m = bench.build_net(dev, confs=c)
        m.mlp_mode = _lib.MLP_TC_FAST
        sc = synthetic.make_scene(5, 1080, 1920, base, seed=20, device=dev)
        torch.cuda.synchronize()
        ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
        stats = ps.stats()
        near, far = sc.near, sc.far
        rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
        sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.matching_volume = [], [], [], None
        del sc
        torch.cuda.empty_cache()
        n = rays_o.shape[0]
        torch.manual_seed(1)
        t_all = m.draw_chunk_randoms(n)
        r0, r1 = sdist.shard_rays(n, rank, world)
        o, d, t = rays_o[r0:r1].contiguous(), rays_d[r0:r1].contiguous(), t_all[r0:r1].to(dev)
        gather = sdist.ImageGather(n, dev) if world > 1 else None

        def step():
            res = m.render_image(ps, o, d, near, far, t_rand=t)
            return gather(res) if gather is not None else res
        for _ in range(max(3, args.warmup)):
            step()
        l0 = _lib.launch_count()
        ms = timed(step, args.steps, world, dev)
        launches = _lib.launch_count() - l0
        line = {"metric": "rays_per_sec", "value": n * args.steps / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "dtype": "f16 MMA, fp32 accumulate (opt-in 1e-2 mode)", "data": "synthetic",
                "config": {"workload": "cfgE: 1080x1920, 5 views (4 src), 128 samples/ray [64,32,16,16], volumes %d->%d"
                                       % (base, base * 8),
                           "rays_per_image": n,
                           "parallelism": "one image, rays sharded x%d + NCCL all-gather" % world if world > 1 else "1 GPU"},
                "gpu_launches": int(launches), "scene_bytes": stats,
                "device_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
