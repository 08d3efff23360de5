#!/usr/bin/env python
"""bench.py — throughput of the SuRF render hot path on B200 (contract: see DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

Workload (BASELINE.json configs[1]): full-image render 576x800, 3 views (2 source views), fp32,
S = 136 samples per ray, synthetic DTU-shaped scene of SURVEY.md §8d (volumes 88 -> 704).  One step =
one full image = 460 800 rays through sampler -> mask -> SDF MLP (+gradient) -> projection gather ->
blending MLP -> compositing.

Under torchrun (N > 1) the default is the north-star multi-GPU workload: ONE image, rays sharded over the
ranks on 256-ray chunk boundaries (surf_b200.dist.shard_rays) and the rendered tiles all-gathered over NCCL
INSIDE the timed region (`scaling: "strong"`); the 512^3 SDF grid is x-slab sharded and all-gathered the same
way.  `--scaling weak` (one image per rank, no data-path collective) is also measured as a secondary key.

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
FLOP_BLEND_PER_POINT_VIEW = 19856  # blending MLP per (point, view)
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
    ap.add_argument("--level0", default="frustum", choices=["frustum", "all"],
                    help="coarsest voxel mask of the synthetic scene: SURVEY 8d's formula (default) or every voxel occupied "
                         "(8d's voxel counts; a real scene with cameras all around)")
    ap.add_argument("--color-path", type=int, default=0, choices=[0, 1, 2],
                    help="A/B: 0 gather then blend (default), 1 gather beside the SDF kernel, 2 gather fused into the blend")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = one image ray-sharded + NCCL all-gather (default); weak = one image per rank")
    ap.add_argument("--mlp-mode", type=int, default=1, choices=[0, 1, 4],
                    help="0 = fp32 FFMA MLP kernels, 1 = tcgen05 kernels with the fp16 hi/lo 3-MMA split (fp32-grade, "
                         "the drop-in default), 4 = tcgen05 single fp16 MMA (opt-in 1e-2 mode)")
    ap.add_argument("--reference-kind", default="auto", choices=["auto", "reference", "port"],
                    help="--impl reference: the unmodified reference when its tree is reachable ($SURF_REF, "
                         "baseline/_ref, /root/reference), else the oracle port")
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


def build_net(device=None, seed=0, confs=None):
    """Random-init network (geometric init => SDF ~ sphere r=0.5) with the feature / PE-frequency
    columns, which the geometric init zeroes, given small random weights so that the sparse-volume path
    contributes to SDF and gradient exactly as in a trained model.  Deterministic."""
    from surf_b200 import conf
    from surf_b200.modules.implicit_surface import ImplicitSurface
    torch.manual_seed(seed)
    m = ImplicitSurface(confs if confs is not None else conf.default_implicit_surface_conf())
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


def cpu_render_sample(O, onet, sc_cpu, rays_o, rays_d, t_rand, stages=False):
    n = rays_o.shape[0]
    near, far = sc_cpu.near.expand(n, 1), sc_cpu.far.expand(n, 1)
    return O.render(onet, rays_o, rays_d, near, far, sc_cpu.matching_volume, sc_cpu.volumes, sc_cpu.sparse_idxes,
                    sc_cpu.mask_volumes, sc_cpu.imgs, sc_cpu.features, sc_cpu.intrs, sc_cpu.c2ws, 1.0, t_rand=t_rand,
                    pts_random=torch.zeros(1, 3), return_stages=stages)


def cpu_sample_rays(args, sc_cpu):
    from surf_b200 import synthetic
    o, d = synthetic.random_pixel_rays(sc_cpu, args.cpu_rays, seed=2)
    t_rand = torch.rand(args.cpu_rays, 4, generator=torch.Generator().manual_seed(0))
    return o, d, t_rand


def cpu_baseline(args, m, sc_cpu, n_runs=2):
    """The reference's CPU path (oracle port, oracle/surf_oracle.py: same ATen ops as the reference) on a
    bounded sample of the same workload: `cpu_rays` random rays of the same image."""
    O, onet = oracle_net(m)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o, d, t_rand = cpu_sample_rays(args, sc_cpu)
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


def parity_block(args, m, ps, sc_cpu, dev):
    """GPU result vs the CPU oracle on the cpu_baseline rays, same jitter, in the benchmarked kernel family: mask
    census, raw scale-relative errors, and the errors normalised by the test tolerance (1e-4 of scale + 1.5 x the
    first-order image of the measured per-point deviations, oracle/surf_oracle.py:composite_envelope)."""
    O, onet = oracle_net(m)
    o, d, t_rand = cpu_sample_rays(args, sc_cpu)
    n = o.shape[0]
    ref = cpu_render_sample(O, onet, sc_cpu, o, d, t_rand, stages=True)
    near, far = sc_cpu.near.expand(n, 1).to(dev), sc_cpu.far.expand(n, 1).to(dev)
    out = m.render(o.to(dev), d.to(dev), near, far, ps, None, None, None, None, None, None, None, None, 1.0, None,
                   t_rand=t_rand, pts_random=torch.zeros(1, 3), return_stages=True)
    B, S = ref["mid_z_vals"].shape
    vm_g = (out["_point_flags"].cpu() & 1).bool().reshape(B, S)
    vm_r = ref["_voxel_mask"].reshape(B, S)
    mism = vm_g != vm_r
    on_boundary = None
    if bool(mism.any()):
        pts = (o[:, None, :] + d[:, None, :] * ref["mid_z_vals"][..., None])[mism]
        dz = (out["mid_z_vals"].cpu().double() - ref["mid_z_vals"].double()).abs()[mism]
        on_boundary = int((O.voxel_round_distance(pts, sc_cpu.mask_volumes) <= 2.0 * dz + 2.5e-7).sum())
    bits = out["_point_views"].cpu()
    vmr = ref["_view_mask"]
    got_vm = torch.stack([(bits >> v) & 1 for v in range(vmr.shape[1])], dim=1).bool()
    view_mism = (got_vm != vmr).any(dim=1) & ref["_compute_mask"]
    same_cross = out["_prev_idx"].cpu().long() == ref["_prev_idx"][:, 0]
    # a sample within rounding noise of a voxel FACE has no robust gradient in the reference itself (the trilinear
    # feature lookup is continuous there, its derivative is not): every sample whose gradient deviates by more than
    # 1e-4 of scale must be proven to sit on a face (tests/helpers.explain_gradient_mismatches); its ray is counted,
    # not compared
    cm = ref["_compute_mask"]
    g_dev = (out["gradients"].cpu().reshape(-1, 3).double() - ref["_grad"].double()).abs().max(dim=1)[0]
    g_scale = max(float(ref["_grad"][cm].abs().max()), 1e-30)
    # ... and inside a voxel the reference's own gradient moves with the last bit of the sample position:
    # tolerance = 1e-4 of scale + 2 x oracle.gradient_position_envelope (first-order image of the measured difference
    # of the two paths' sample depths + one ulp of a grid coordinate)
    g_env = O.gradient_position_envelope(onet, ref, out["mid_z_vals"], o, d, sc_cpu.volumes, sc_cpu.sparse_idxes)
    g_tol = 1e-4 * g_scale + 2.0 * g_env
    g_cand = (g_dev > 1e-4 * g_scale) & cm
    g_bad = torch.zeros_like(g_cand)            # deviating samples ON a voxel face: ray counted, not compared
    g_unexplained = 0
    if bool(g_cand.any()):
        dz_b = (out["mid_z_vals"].cpu().double() - ref["mid_z_vals"].double()).abs().reshape(-1)[g_cand]
        dn_b = d.double().norm(dim=1)[:, None].expand(B, S).reshape(-1)[g_cand]
        face = O.voxel_face_margin(ref["_pts"][g_cand], sc_cpu.sparse_idxes, 2.0 * dz_b * dn_b) < 2.5
        g_bad[g_cand.nonzero()[:, 0][face]] = True
        g_unexplained = int((~face & (g_dev[g_cand] > g_tol[g_cand])).sum())
    g_on_face = int(g_bad.sum())
    rows = ~mism.any(dim=1) & ~view_mism.reshape(B, S).any(dim=1) & same_cross & ~g_bad.reshape(B, S).any(dim=1)
    keep_p = rows[:, None].expand(B, S).reshape(-1) & ref["_compute_mask"]
    sdf_g = out["sparse_sdf"][-B * S:].cpu()
    grad_g = out["gradients"].cpu().reshape(-1, 3)
    col_g = out["_point_color"].cpu().reshape(-1, 3)

    def rel(a, b):
        a, b = a.double(), b.double()
        return float((a - b).abs().max() / max(float(b.abs().max()), 1e-2)) if b.numel() else 0.0

    inv_s = torch.exp(onet.variance * 10.0).clip(1e-6, 1e6)
    rot = torch.inverse(sc_cpu.c2ws[0, :3, :3])
    env, _ = O.composite_envelope(ref, sdf_g, grad_g, col_g, o, d, inv_s, rot)
    max_rel = {"sdf": rel(sdf_g[keep_p], ref["_sdf"][keep_p]), "gradient": rel(grad_g[keep_p], ref["_grad"][keep_p]),
               "point_color": rel(col_g[keep_p], ref["_color"].reshape(-1, 3)[keep_p])}
    err_over_tol = {}
    for k in ("color_fine", "render_depth", "sdf_depth", "normal"):
        a = out[k].cpu().double().reshape(B, -1)[rows]
        b = ref[k].double().reshape(B, -1)[rows]
        max_rel[k] = rel(a, b)
        scale = max(float(b.abs().max()), 1e-2)
        err_over_tol[k] = float(((a - b).abs() / (1e-4 * scale + 1.5 * env[k][rows])).max())
    g_ok = cm & ~g_bad
    err_over_tol["gradient"] = float((g_dev[g_ok] / g_tol[g_ok]).max()) if bool(g_ok.any()) else 0.0
    within = (max_rel["sdf"] <= 1e-4 and all(v <= 1.0 for v in err_over_tol.values())
              and (on_boundary is None or on_boundary == int(mism.sum())) and g_unexplained == 0)
    return {"mode": int(m.mlp_mode), "rays": int(B), "rays_compared": int(rows.sum()),
            "mask_mismatches": int(mism.sum()), "mask_mismatches_on_voxel_boundary": on_boundary,
            "view_mask_mismatches": int(view_mism.sum()), "crossing_index_differs": int((~same_cross).sum()),
            "gradient_deviations_beyond_1e-4": int(g_cand.sum()), "of_which_on_a_voxel_face": g_on_face,
            "of_which_beyond_the_position_envelope_and_not_on_a_face": g_unexplained,
            "max_rel": max_rel,"err_over_tolerance": err_over_tol, "within_tolerance": bool(within),
            "tolerance": "per-point sdf 1e-4 of scale; gradient 1e-4 of scale + 2 x sample-position conditioning envelope, "
                         "deviations beyond it proven on a voxel face; composited 1e-4 of scale + 1.5 x first-order image of "
                         "the measured per-point deviations (inv_s amplification, DESIGN.md §2)"}


def config_dict(args, world, strong):
    return {"workload": "full-image render %dx%d, %d views (%d src), S=%d samples/ray, volumes %d->%d, fp32 (configs[1])"
            % (args.height, args.width, args.views, args.views - 1, S_TOTAL, args.base, args.base * 8),
            "rays_per_image": args.height * args.width,
            "parallelism": ("1 GPU" if world == 1 else
                            ("one image, rays sharded x%d on 256-ray chunk boundaries + NCCL all-gather of tiles" % world
                             if strong else "%d independent images (one per rank), no collective" % world)),
            "l2": "inputs larger than L2 (prepared scene > 2 GB, ~13 GB of per-step intermediates)",
            **({"level0": "all voxels of the coarsest level occupied"} if getattr(args, "level0", "frustum") == "all" else {})}


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores — the UNMODIFIED
    reference (its stock ImplicitSurface.render, 256-ray chunks as validate() issues them) when its source tree is
    reachable, else the oracle port.  Rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from surf_b200 import synthetic
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sc = synthetic.make_scene(args.views, args.height, args.width, args.base, seed=1, device=dev, level0=args.level0)
    sc_cpu = sc.to("cpu")
    del sc
    m = build_net()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.cpu_rays
    o, d = synthetic.random_pixel_rays(sc_cpu, n, seed=2)
    g = torch.Generator().manual_seed(0)

    kind = "port"
    ref_net = None
    if args.reference_kind in ("auto", "reference"):
        try:
            import ref_loader
            if ref_loader.find_reference() is not None:
                IS = ref_loader.load_reference()
                from surf_b200.conf import default_implicit_surface_conf
                ref_net = IS.ImplicitSurface(default_implicit_surface_conf())
                ref_net.load_state_dict(m.state_dict(), strict=True)
                ref_net.eval()
                kind = "reference"
        except Exception as e:      # the reference tree is not importable here: fall back to the port
            print("reference tree not usable (%s): timing the oracle port" % e, file=sys.stderr)
            ref_net = None
    if args.reference_kind == "reference" and ref_net is None:
        emit({"impl": "reference", "unavailable": "reference source tree not reachable on this box"})
        return

    def step_port(oo, dd, tt):
        O, onet = oracle_net(m)
        return cpu_render_sample(O, onet, sc_cpu, oo, dd, tt)

    def step_reference(oo, dd, tt):
        # stock path: validate() calls render() per 256-ray chunk (implicit_surface.py:367-385); the reference
        # draws its own jitter from the global generator
        nn_ = oo.shape[0]
        near, far = sc_cpu.near.expand(nn_, 1), sc_cpu.far.expand(nn_, 1)
        with torch.no_grad():
            for a in range(0, nn_, 256):
                b = min(nn_, a + 256)
                ref_net.render(oo[a:b], dd[a:b], near[a:b], far[a:b], sc_cpu.matching_volume, sc_cpu.volumes,
                               sc_cpu.sparse_idxes, sc_cpu.mask_volumes, sc_cpu.imgs, sc_cpu.features,
                               sc_cpu.features, sc_cpu.intrs, sc_cpu.c2ws, 1.0, None)

    step = step_reference if ref_net is not None else step_port
    if ref_net is not None:
        n = min(n, 2048)            # the stock chunked path is dispatch-bound: keep the sample to ~10-30 s
        o, d = o[:n], d[:n]
    for _ in range(max(1, min(args.warmup, 2))):
        step(o[:256], d[:256], torch.rand(256, 4, generator=g))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(o, d, torch.rand(n, 4, generator=g))
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    port = None
    if ref_net is not None:         # the port side by side
        step_port(o[:256], d[:256], torch.rand(256, 4, generator=g))
        t1 = time.perf_counter()
        step_port(o, d, torch.rand(n, 4, generator=g))
        port = {"value": n / (time.perf_counter() - t1), "unit": "rays/s", "kind": "port"}
    sample = ("%d rays per step (bounded sample of the %d-ray image), %s" % (
        n, args.height * args.width,
        "UNMODIFIED reference ImplicitSurface.render in 256-ray chunks" if ref_net is not None
        else "oracle port of the reference render() (reference tree not on this box)"))
    line = {
        "impl": "reference", "metric": "rays_per_sec", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if (args.gpus > 1 and args.scaling == "strong") else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, args.gpus, args.gpus > 1 and args.scaling == "strong"),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if port is not None:
        line["cpu_baseline_port"] = port
    emit(line)


# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    rank, world, local = dist_env()
    import torch.distributed as dist
    from surf_b200 import _lib, synthetic
    from surf_b200 import dist as sdist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (surf_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    strong = args.scaling == "strong" and world > 1

    # ---- scene + network (generated on the device; scene_prepare timed separately) ----------------
    # every rank holds the same scene (seed 1): strong scaling shards ONE image; the secondary weak measurement
    # renders the same image on every rank
    sc = synthetic.make_scene(args.views, args.height, args.width, args.base, seed=1, device=dev, level0=args.level0)
    m = build_net(dev)
    m.mlp_mode = args.mlp_mode
    m.color_path = int(args.color_path)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    torch.cuda.synchronize()
    prepare_s = time.perf_counter() - t0
    stats = ps.stats()
    near, far = sc.near, sc.far
    do_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline)
    sc_cpu = sc.to("cpu") if do_cpu else None
    voxels = sc.voxel_counts()
    rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
    # free the reference-layout tensors: the hot path only needs the prepared scene
    sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.matching_volume = [], [], [], None
    del sc
    torch.cuda.empty_cache()

    n_rays = rays_o.shape[0]
    torch.manual_seed(1234)
    t_rand_all = m.draw_chunk_randoms(n_rays)

    def shard(strong_):
        r0, r1 = sdist.shard_rays(n_rays, rank, world) if strong_ else (0, n_rays)
        return (rays_o[r0:r1].contiguous(), rays_d[r0:r1].contiguous(), t_rand_all[r0:r1].to(dev), r0, r1)

    my_o, my_d, my_t, r0, r1 = shard(strong)
    my_n = r1 - r0
    gather = sdist.ImageGather(n_rays, dev) if strong else None

    def step_device():
        res = m.render_image(ps, my_o, my_d, near, far, t_rand=my_t)
        if gather is not None:
            return gather(res)              # NCCL all-gather of the packed (R, 8) records, inside the timed region
        return res

    # host-buffer (e2e) arm: pinned host rays in, results back to pinned host, jitter drawn on the host
    h_o, h_d = my_o.cpu().pin_memory(), my_d.cpu().pin_memory()
    out_n = n_rays if strong else my_n
    out_host = {"color_fine": torch.empty((out_n, 3)).pin_memory(), "val_normal": torch.empty((out_n, 3)).pin_memory(),
                "sdf_depth": torch.empty((out_n, 1)).pin_memory(), "render_depth": torch.empty((out_n,)).pin_memory()}
    # per step, whole job: rays (o, d) + the jitter table of every rank in; the image out (rank 0 reads the
    # gathered image under strong scaling, every rank its own image under weak scaling)
    n_job = n_rays if strong else n_rays * world
    h2d = n_job * (3 + 3 + 4) * 4
    d2h = (n_rays if strong else n_rays * world) * 8 * 4

    # strong scaling: the reference's jitter stream is sequential over the image (Q1), so every rank has to consume the
    # draws of the other ranks' chunks too (~20 ms of mt19937 per image, more than the 8-GPU device time): a worker
    # thread draws the table of image k+1 while image k renders (surf_b200/dist.py: JitterPrefetcher)
    jitter = None      # created right before the e2e loop (the worker owns the global generator while it is open)

    def step_e2e():
        o = h_o.to(dev, non_blocking=True)
        d = h_d.to(dev, non_blocking=True)
        if strong:
            res = m.render_image(ps, o, d, near, far, t_rand=jitter.next())
            res = gather(res)
            if rank == 0:
                for k, v in out_host.items():
                    v.copy_(res[k], non_blocking=True)
        else:
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
        m.render_image(ps, my_o, my_d, near, far, t_rand=my_t)
    torch.cuda.synchronize()
    kt = _lib.timing_read()
    _lib.timing_enable(False)
    # valid (evaluated) sample points of one step, from the flags of an untimed pass
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
    kernel_ms = {k: v[0] / n_prof for k, v in kt.items() if v[1] > 0}
    prof = {}
    prof_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof_json):
        try:
            prof = json.load(open(prof_json))
        except Exception:
            prof = {}
    traffic = prof.get("sdf_mlp_grad", {}).get("dram_bytes_per_launch")
    tc = args.mlp_mode >= 1
    V = args.views - 1
    kname = ("k_sdf_tc2<GRAD=true> (tcgen05: sparse gather + SDF MLP forward + input gradient, fp16 hi/lo 3-MMA)" if tc
             else "k_sdf_mlp<GRAD=true> (fp32 FFMA: sparse gather + SDF MLP forward + input gradient)")
    # the other kernels of the step against their own roofline (algorithmic work per SURVEY §8d / DESIGN §4)
    P_all = my_n * S_TOTAL
    others = {}

    def add(kind, name, bound, work, unit_scale, peak):
        if kind in kernel_ms and kernel_ms[kind] > 0:
            ach = work / (kernel_ms[kind] * 1e-3) / unit_scale
            others[name] = {"bound": bound, "ms": kernel_ms[kind], "achieved": ach, "peak": peak,
                            "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": ach / peak}

    fused = tc and V in (2, 4) and args.color_path == 2
    add("blend", ("k_blend_tm<FUSED> (projection gather + blending MLP in one kernel)" if fused else
                  "k_blend_tm" if tc and V in (2, 4) else "k_blend_tc" if tc else "k_blend"), "tensor", n_eval * V * FLOP_BLEND_PER_POINT_VIEW, 1e12, pk["bf16_tflops"])
    add("lookup_feature", "k_lookup_feature", "hbm", n_eval * V * 304.0, 1e9, pk["hbm_gbs"])
    add("sample_rays", "k_sample_rays", "hbm", my_n * 8192.0, 1e9, pk["hbm_gbs"])
    add("point_flags", "k_point_flags", "hbm", P_all * 21.0, 1e9, pk["hbm_gbs"])
    add("composite", "k_composite", "hbm", P_all * 33.0 + my_n * 36.0, 1e9, pk["hbm_gbs"])
    roofline = {"kernel": kname, "bound": "tensor",
                "achieved": achieved_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved_tflops / pk["bf16_tflops"], "traffic": traffic, "peak_source": pk_src,
                "frac_of_sustained_peak": achieved_tflops / pk["bf16_tflops_sustained"],
                "algorithmic": "%d evaluated points/step x %d FLOP (fwd+input-grad, SURVEY 8d) over %d launches/step"
                               % (n_eval, FLOP_PER_POINT_FWD_BWD, mlp_launches // n_prof),
                "note": ("achieved counts ALGORITHMIC flops (one fp32 product each); the fp32-parity mode issues every "
                         "product as 3 fp16 MMAs (hi*hi + lo*hi + hi*lo), so the tensor pipe executes 3x this figure"
                         if tc else "fp32 FFMA edition of the MLP: fraction is against the measured bf16 tensor peak"),
                "tensor_work_tflops": achieved_tflops * (3.0 if tc else 0.0),
                "kernel_ms_per_step": kernel_ms, "other_kernels": others}

    # ---- e2e through the public API with host buffers -------------------------------------------------
    if strong:
        torch.manual_seed(1234)
        jitter = sdist.JitterPrefetcher(m, n_rays, r0, r1)
    for _ in range(2):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e = timed(step_e2e, e2e_steps)
    if jitter is not None:
        jitter.close()
    e2e_value = total_rays * e2e_steps / (ms_e2e * 1e-3)

    # ---- secondary: the other scaling mode -------------------------------------------------------------
    other_scaling = None
    if world > 1:
        o2, d2, t2, a0, a1 = shard(not strong)
        g2 = sdist.ImageGather(n_rays, dev) if not strong else None

        def step_other():
            r = m.render_image(ps, o2, d2, near, far, t_rand=t2)
            return g2(r) if g2 is not None else r
        for _ in range(2):
            step_other()
        ms_o = timed(step_other, 2)
        tot = n_rays if not strong else n_rays * world
        other_scaling = {"scaling": "weak" if strong else "strong", "value": tot * 2 / (ms_o * 1e-3), "unit": "rays/s",
                         "ms_per_step": ms_o / 2,
                         "what": ("one full image per rank, no collective" if strong
                                  else "one image ray-sharded + NCCL all-gather")}

    # ---- SDF grid (configs[2]): 512^3 dense query, x-slab sharded across ranks + all-gather of the slabs ----
    grid = None
    if args.grid > 0:
        R = args.grid
        x0, x1 = sdist.shard_planes(R, rank, world)

        def step_grid():
            u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], R, x_range=(x0, x1))
            if world > 1:
                u = sdist.gather_grid(u, R)       # every rank ends up with the full grid (537 MB at 512^3)
            return u
        # warm-up: the outputs must come from the caching allocator, not from a cudaMalloc inside the timed region
        step_grid()
        step_grid()
        ms_grid = timed(step_grid, 2) / 2
        grid = {"value": R ** 3 / (ms_grid * 1e-3), "unit": "pts/s", "resolution": R, "ms": ms_grid,
                "mode": "dense (parity mode, Q16)", "tflops": R ** 3 * FLOP_PER_POINT_FWD / (ms_grid * 1e-3) / 1e12,
                "parallelism": "1 GPU" if world == 1 else "x-slabs over %d ranks + NCCL all-gather of the slabs "
                               "inside the timed region" % world}
        # the rest of extract_geometry on the gathered grid: GPU marching cubes (replaces host PyMCubes), and the
        # opt-in sparsified query (BASELINE configs[2] "with surface-region sparsification": points outside the 4-level
        # voxel mask get a constant — NOT result-identical outside the mask, Q16)
        from surf_b200 import mesh as smesh
        u_full = step_grid()

        def step_mc():
            return smesh.marching_cubes_device(u_full, 0.0)
        v_mc, t_mc = step_mc()
        ms_mc = timed(step_mc, 2) / 2
        grid["marching_cubes"] = {"ms": ms_mc, "vertices": int(v_mc.shape[0]), "triangles": int(t_mc.shape[0]),
                                  "gb_per_s": (R ** 3 * 12.0) / (ms_mc * 1e-3) / 1e9,
                                  "note": "GPU marching cubes of the full grid incl. the host read of the output sizes; "
                                          "12 B / grid point algorithmic (4 B read + 4 B code word written and re-read)"}
        del u_full, v_mc, t_mc

        def step_grid_sparse():
            u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], R, x_range=(x0, x1), sparsify=True)
            if world > 1:
                u = sdist.gather_grid(u, R)
            return u
        step_grid_sparse()
        ms_sp = timed(step_grid_sparse, 2) / 2
        grid["sparsified"] = {"value": R ** 3 / (ms_sp * 1e-3), "unit": "pts/s", "ms": ms_sp,
                              "mode": "opt-in: SDF evaluated only inside the voxel mask (Q16)"}
        del step_grid, step_grid_sparse
        torch.cuda.empty_cache()

    # ---- opt-in reduced-precision mode (north_star: 1e-2 mode), reported next to the headline, not as it ----
    fast = None
    if args.mlp_mode == 1:
        m.mlp_mode = _lib.MLP_TC_FAST
        for _ in range(2):
            step_device()
        ms_fast = timed(step_device, 2)
        fast = {"value": total_rays * 2 / (ms_fast * 1e-3), "unit": "rays/s", "ms_per_step": ms_fast / 2,
                "mode": "tcgen05, one fp16 MMA per product (tolerance 1e-2; measured error ~1e-3)"}
        m.mlp_mode = args.mlp_mode

    cpu = parity = None
    if do_cpu:
        cpu = cpu_baseline(args, m, sc_cpu)
        parity = parity_block(args, m, ps, sc_cpu, dev)

    if rank == 0:
        cfg = config_dict(args, world, strong)
        line = {
            "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32 (tcgen05 fp16 hi/lo split, fp32 accumulate)" if tc else "f32", "data": "synthetic",
            "config": cfg,
            "detail": {"voxels_fine_to_coarse": voxels, "evaluated_points_per_step_per_rank": n_eval,
                       "rays_per_step_per_rank": my_n, "mlp_mode": args.mlp_mode, "scene_prepare_s": prepare_s,
                       "scene_bytes": stats},
            "clocks": clock_info, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / e2e_steps,
                    "api": "ImplicitSurface.render_image%s: pinned host rays -> device, reference-order jitter drawn on "
                           "the host%s, results -> pinned host"
                           % ((" + surf_b200.dist.ImageGather (NCCL)", " (whole-image stream per rank, one image ahead on a "
                               "worker thread: dist.JitterPrefetcher)") if strong else ("", ""))},
            "roofline": roofline,
            "sdf_grid": grid,
            "reduced_precision_mode": fast,
        }
        if other_scaling is not None:
            line["secondary_scaling"] = other_scaling
        if cpu is not None:
            line["cpu_baseline"] = cpu
            line["parity"] = parity
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the process' original stdout; everything else (NCCL's version banner, library chatter)
    was redirected to stderr in main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                      # C-level writes to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
