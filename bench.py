#!/usr/bin/env python
"""bench.py — throughput of the SuRF render hot path on B200 (contract: see DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): full-image render 576x800, 3 views (2 source views), fp32,
S = 136 samples per ray, synthetic DTU-shaped scene of SURVEY.md §8d (volumes 88 -> 704).  One step =
one full image = 460 800 rays through sampler -> mask -> SDF MLP (+gradient) -> projection gather ->
blending MLP -> compositing.  Under torchrun (N > 1) every rank renders one full image of its own
(weak scaling; rays are independent, there is no data-path collective), `value` = all rays / max time.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import platform
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W, NV, BASE = 576, 800, 3, 88
S_TOTAL = 136
GRID_RES = 512
FLOP_PER_POINT_FWD = 198480        # SDF MLP forward, sdf-only head
FLOP_PER_POINT_FWD_BWD = 396960    # forward + input-gradient (reverse pass = same MACs)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}   # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="surf_b200", choices=["surf_b200", "reference"])
    ap.add_argument("--base", type=int, default=BASE, help="coarsest volume dim (88 = confs/surf.conf)")
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--views", type=int, default=NV)
    ap.add_argument("--grid", type=int, default=GRID_RES, help="SDF grid resolution (0 = skip)")
    ap.add_argument("--cpu-rays", type=int, default=8192, help="rays per CPU-baseline sample / reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--mlp-mode", type=int, default=1, choices=[0, 1, 3, 4, 5],
                    help="0 = fp32 FFMA MLP kernels, 1 = tcgen05 kernels with the fp16 hi/lo 3-MMA split (fp32-grade), "
                         "4 = tcgen05 single fp16 MMA (opt-in 1e-2 mode)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"]))}, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS, bf16_tflops_sustained=1400.0), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_net(device=None, seed=0):
    """Random-init network (geometric init => SDF ~ sphere r=0.5) with the feature / PE-frequency
    columns, which the geometric init zeroes, given small random weights so that the sparse-volume path
    contributes to SDF and gradient exactly as in a trained model.  Deterministic."""
    from surf_b200 import conf
    from surf_b200.modules.implicit_surface import ImplicitSurface
    torch.manual_seed(seed)
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("weight_v"):
                zero_cols = (p == 0).to(p.dtype)
                p.add_(torch.randn(p.shape, generator=g) * 0.02 * zero_cols)
    return m if device is None else m.to(device)


def oracle_net(m):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import surf_oracle as O
    return O, O.OracleNet({k: v.detach().cpu() for k, v in m.state_dict().items()})


def cpu_render_sample(O, onet, sc_cpu, rays_o, rays_d, t_rand):
    n = rays_o.shape[0]
    near, far = sc_cpu.near.expand(n, 1), sc_cpu.far.expand(n, 1)
    return O.render(onet, rays_o, rays_d, near, far, sc_cpu.matching_volume, sc_cpu.volumes, sc_cpu.sparse_idxes,
                    sc_cpu.mask_volumes, sc_cpu.imgs, sc_cpu.features, sc_cpu.intrs, sc_cpu.c2ws, 1.0, t_rand=t_rand,
                    pts_random=torch.zeros(1, 3))


def cpu_baseline(args, m, sc_cpu, n_runs=2):
    """The reference's CPU path (oracle port, oracle/surf_oracle.py: same ATen ops as the reference) on a
    bounded sample of the same workload: `cpu_rays` random rays of the same image."""
    from surf_b200 import synthetic
    O, onet = oracle_net(m)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o, d = synthetic.random_pixel_rays(sc_cpu, args.cpu_rays, seed=2)
    t_rand = torch.rand(args.cpu_rays, 4, generator=torch.Generator().manual_seed(0))
    cpu_render_sample(O, onet, sc_cpu, o[:64], d[:64], t_rand[:64])          # warm-up
    ts = []
    for _ in range(n_runs):
        t0 = time.perf_counter()
        cpu_render_sample(O, onet, sc_cpu, o, d, t_rand)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    t = ts[len(ts) // 2] if len(ts) % 2 else 0.5 * (ts[len(ts) // 2 - 1] + ts[len(ts) // 2])
    return {"value": args.cpu_rays / t, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d random rays of the same %dx%d image x %d samples, oracle/surf_oracle.py render(), torch %d "
                      "threads, median of %d" % (args.cpu_rays, args.height, args.width, S_TOTAL, cores, n_runs),
            "cpu": platform.processor() or platform.machine(), "seconds": t}


def config_dict(args, world, extra=None):
    c = {"workload": "full-image render %dx%d, %d views (%d src), S=%d samples/ray, volumes %d->%d, fp32 (configs[1])"
         % (args.height, args.width, args.views, args.views - 1, S_TOTAL, args.base, args.base * 8),
         "rays_per_step_per_gpu": args.height * args.width, "parallelism": "ray-sharded x%d" % world,
         "l2": "inputs larger than L2 (prepared scene > 2 GB, ~13 GB of per-step intermediates)"}
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from surf_b200 import synthetic
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sc = synthetic.make_scene(args.views, args.height, args.width, args.base, seed=1, device=dev)
    sc_cpu = sc.to("cpu")
    del sc
    m = build_net()
    O, onet = oracle_net(m)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.cpu_rays
    o, d = synthetic.random_pixel_rays(sc_cpu, n, seed=2)
    g = torch.Generator().manual_seed(0)
    for _ in range(max(1, args.warmup)):
        cpu_render_sample(O, onet, sc_cpu, o[:128], d[:128], torch.rand(128, 4, generator=g))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_render_sample(O, onet, sc_cpu, o, d, torch.rand(n, 4, generator=g))
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = "%d rays per step (bounded sample of the %d-ray image), oracle port of the reference render()" % (
        n, args.height * args.width)
    print(json.dumps({
        "impl": "reference", "metric": "rays_per_sec", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, 1, {"rays_per_step": n}),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    rank, world, local = dist_env()
    import torch.distributed as dist
    from surf_b200 import _lib, synthetic
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (surf_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    _lib.set_mlp_mode(args.mlp_mode)

    # ---- scene + network (generated on the device; scene_prepare timed separately) ----------------
    # weak scaling: rank r renders the image of "scene r" (same shape, different seed)
    strong = args.scaling == "strong" and world > 1
    sc = synthetic.make_scene(args.views, args.height, args.width, args.base, seed=1 if strong else 1 + rank, device=dev)
    m = build_net(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    torch.cuda.synchronize()
    prepare_s = time.perf_counter() - t0
    stats = ps.stats()
    near, far = sc.near, sc.far
    intrs_h, c2ws = sc.intrs, sc.c2ws
    do_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline)
    sc_cpu = sc.to("cpu") if do_cpu else None
    voxels = sc.voxel_counts()
    rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
    # free the reference-layout tensors: the hot path only needs the prepared scene
    sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.matching_volume = [], [], [], None
    del sc
    torch.cuda.empty_cache()

    n_rays = rays_o.shape[0]
    if strong:
        per = ((n_rays + world - 1) // world + 255) // 256 * 256
        r0, r1 = min(n_rays, rank * per), min(n_rays, (rank + 1) * per)
    else:
        r0, r1 = 0, n_rays
    torch.manual_seed(1234)
    t_rand_all = m.draw_chunk_randoms(n_rays)
    my_o, my_d = rays_o[r0:r1].contiguous(), rays_d[r0:r1].contiguous()
    my_t = t_rand_all[r0:r1].to(dev)
    my_n = r1 - r0

    def step_device():
        return m.render_image(ps, my_o, my_d, near, far, t_rand=my_t)

    # host-buffer (e2e) arm: pinned host rays in, results back to pinned host, jitter drawn on the host
    h_o, h_d = my_o.cpu().pin_memory(), my_d.cpu().pin_memory()
    out_host = {"color_fine": torch.empty((my_n, 3)).pin_memory(), "val_normal": torch.empty((my_n, 3)).pin_memory(),
                "sdf_depth": torch.empty((my_n, 1)).pin_memory(), "render_depth": torch.empty((my_n,)).pin_memory()}
    h2d = h_o.numel() * 4 + h_d.numel() * 4 + my_n * 4 * 4
    d2h = sum(v.numel() * 4 for v in out_host.values())

    def step_e2e():
        o = h_o.to(dev, non_blocking=True)
        d = h_d.to(dev, non_blocking=True)
        res = m.render_image(ps, o, d, near, far)        # jitter drawn on the host inside, batch by batch
        for k, v in out_host.items():
            v.copy_(res[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    # ---- warm-up, then the timed region ---------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        res = step_device()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    _lib.timing_enable(False)
    launches0 = _lib.launch_count()
    ms_total = timed(step_device, args.steps)
    launches = _lib.launch_count() - launches0
    clock_info = clocks.stop() if rank == 0 else None
    total_rays = n_rays if strong else n_rays * world
    value = total_rays * args.steps / (ms_total * 1e-3)

    # ---- per-kernel device time (separate pass, events around each launch on its own stream) ---------
    _lib.timing_enable(True)
    _lib.timing_read()
    n_prof = max(1, min(2, args.steps))
    for _ in range(n_prof):
        step_device()
    torch.cuda.synchronize()
    kt = _lib.timing_read()
    _lib.timing_enable(False)
    # valid (evaluated) sample points of one step, from the flags of an untimed pass
    lib = _lib.load()
    n_eval = 0
    with torch.no_grad():
        step = max(256, (m.ray_batch // 256) * 256)
        for a in range(0, my_n, step):
            b = min(my_n, a + step)
            t = m._render_device(ps, my_o[a:b], my_d[a:b], near.expand(b - a, 1), far.expand(b - a, 1), my_t[a:b], None,
                                 1.0, 256, stages=True, lean=True)
            n_eval += int(((t["point_flags"] >> 1) & 1).sum())
    mlp_ms, mlp_launches = kt["sdf_mlp_grad"]
    pk, pk_src = peaks()
    achieved_tflops = (n_eval * n_prof * FLOP_PER_POINT_FWD_BWD) / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    kernel_share = {k: v[0] / n_prof for k, v in kt.items() if v[1] > 0}
    traffic = None
    prof_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof_json):
        try:
            traffic = json.load(open(prof_json)).get("sdf_mlp_grad", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    tc = args.mlp_mode >= 1
    kname = ("k_sdf_tc2<GRAD=true> (tcgen05: sparse gather + SDF MLP forward + input gradient, fp16 hi/lo 3-MMA)" if tc
             else "k_sdf_mlp<GRAD=true> (fp32 FFMA: sparse gather + SDF MLP forward + input gradient)")
    roofline = {"kernel": kname, "bound": "tensor",
                "achieved": achieved_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved_tflops / pk["bf16_tflops"], "traffic": traffic, "peak_source": pk_src,
                "algorithmic": "%d evaluated points/step x %d FLOP (fwd+input-grad, SURVEY 8d) over %d launches/step"
                               % (n_eval, FLOP_PER_POINT_FWD_BWD, mlp_launches // n_prof),
                "note": ("achieved counts ALGORITHMIC flops (one fp32 product each); the fp32-parity mode issues every "
                         "product as 3 fp16 MMAs (hi*hi + lo*hi + hi*lo), so the tensor pipe executes 3x this figure"
                         if tc else "fp32 FFMA edition of the MLP: fraction is against the measured bf16 tensor peak"),
                "tensor_work_tflops": achieved_tflops * (3.0 if tc else 0.0),
                "kernel_ms_per_step": kernel_share}

    # ---- e2e through the public API with host buffers -------------------------------------------------
    for _ in range(2):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_value = total_rays * e2e_steps / (ms_e2e * 1e-3)

    # ---- SDF grid (configs[2]): 512^3 dense query, x-slab sharded across ranks ------------------------
    grid = None
    if args.grid > 0:
        R = args.grid
        planes = (R + world - 1) // world
        x0, x1 = min(R, rank * planes), min(R, (rank + 1) * planes)
        # warm-up over the full range: the 512^3 output (512 MB) must come from the caching allocator, not from a
        # cudaMalloc inside the timed region
        m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], R, x_range=(x0, x1))
        ms_grid = timed(lambda: m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], R, x_range=(x0, x1)), 2) / 2
        grid = {"value": R ** 3 / (ms_grid * 1e-3), "unit": "pts/s", "resolution": R, "ms": ms_grid,
                "mode": "dense (parity mode, Q16)", "tflops": R ** 3 * FLOP_PER_POINT_FWD / (ms_grid * 1e-3) / 1e12}

    # ---- opt-in reduced-precision mode (north_star: 1e-2 mode), reported next to the headline, not as it ----
    fast = None
    if args.mlp_mode == 1:
        _lib.set_mlp_mode(4)
        for _ in range(2):
            step_device()
        ms_fast = timed(step_device, 2)
        fast = {"value": total_rays * 2 / (ms_fast * 1e-3), "unit": "rays/s", "ms_per_step": ms_fast / 2,
                "mode": "tcgen05, one fp16 MMA per product (tolerance 1e-2; measured error ~1e-3)"}
        _lib.set_mlp_mode(args.mlp_mode)

    cpu = None
    if do_cpu:
        cpu = cpu_baseline(args, m, sc_cpu)

    if rank == 0:
        line = {
            "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32 (tcgen05 fp16 hi/lo split, fp32 accumulate)" if tc else "f32", "data": "synthetic",
            "config": config_dict(args, world, {"voxels_fine_to_coarse": voxels, "evaluated_points_per_step": n_eval, "mlp_mode": args.mlp_mode,
                                                "scene_prepare_s": prepare_s, "scene_bytes": stats}),
            "clocks": clock_info, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / e2e_steps,
                    "api": "ImplicitSurface.render_image: pinned host rays -> device, reference-order jitter drawn on "
                           "the host, results -> pinned host"},
            "roofline": roofline,
            "sdf_grid": grid,
            "reduced_precision_mode": fast,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
