"""Timing experiments for the ping-pong SDF kernel (csrc/sdf_tc3.cu) on the bench scene.  Needs a library built with
-DT3_DEBUG (SURF_NVCC_EXTRA=-DT3_DEBUG python -m surf_b200.build --force); results with flags != 0 are INVALID numbers,
only the time is of interest.

    python tools/sdf_bench.py [flag sets ...]      e.g.  0 4 8 16 32 64 128 256 12 124
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                       # noqa: E402

import bench                       # noqa: E402
from surf_b200 import _lib, synthetic   # noqa: E402


def main():
    import time
    t00 = time.time()
    sets = [int(a) for a in sys.argv[1:]] or [0]
    dev = "cuda:0"
    sc = synthetic.make_scene(3, 576, 800, 88, seed=1, device=dev)
    m = bench.build_net(dev)
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    near, far = sc.near, sc.far
    rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
    sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.matching_volume = [], [], [], None
    torch.cuda.empty_cache()
    n = 2 * 65536
    o, d = rays_o[200 * 800:200 * 800 + n].contiguous(), rays_d[200 * 800:200 * 800 + n].contiguous()
    torch.manual_seed(1)
    t = m.draw_chunk_randoms(n).to(dev)
    lib = _lib.load()
    print("setup %.1f s" % (time.time() - t00), flush=True)
    have_dbg = hasattr(lib, "surf_debug_flags")
    for f in sets:
        if f and not have_dbg:
            print("flags %d: library built without -DT3_DEBUG" % f)
            continue
        if have_dbg:
            lib.surf_debug_flags(int(f))
        for _ in range(2):
            m.render_image(ps, o, d, near, far, t_rand=t)
        torch.cuda.synchronize()
        _lib.timing_enable(True)
        _lib.timing_read()
        for _ in range(3):
            m.render_image(ps, o, d, near, far, t_rand=t)
        torch.cuda.synchronize()
        kt = _lib.timing_read()
        _lib.timing_enable(False)
        ms, nl = kt["sdf_mlp_grad"]
        if have_dbg and hasattr(lib, "surf_debug_prof_read"):
            import ctypes
            buf = (ctypes.c_ulonglong * 16)()
            lib.surf_debug_prof_read(buf)
            v = [int(b) for b in buf]
            names = ["issuer total", "  wait a_ready", "  wait weights", "  wait stage_ready", "  wait finish_done",
                     "epilogue w0 total", "  wait d_full", "helper total", "  stage", "  wait tile_done", "  finish",
                     "loader total", "  wait w_empty"]
            print("    CTA 0 clocks (sum over %d launches incl. warm-up): " % 0 + ", ".join(
                "%s %.2fM" % (n_.strip(), x / 1e6) for n_, x in zip(names, v)))
        print("[%.0f s] flags %4d: k_sdf_pp<GRAD> %.3f ms per 65536-ray launch set (%d launches)  -> %.1f ms per 460800-ray image" % (
            time.time() - t00, f, ms / nl, nl, ms / nl * 460800 / 65536), flush=True)
    if have_dbg:
        lib.surf_debug_flags(0)
    # forward-only: the SDF grid
    for f in sets:
        if f and not have_dbg:
            continue
        if have_dbg:
            lib.surf_debug_flags(int(f))
        m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 256)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 256)
        e1.record()
        torch.cuda.synchronize()
        print("flags %4d: 256^3 grid %.3f ms -> %.2f G pts/s" % (f, e0.elapsed_time(e1), 256 ** 3 / e0.elapsed_time(e1) / 1e6))


if __name__ == "__main__":
    main()
