"""Parity AT THE BENCHMARKED CONFIGURATIONS, in the benchmarked kernel family (tcgen05, MLP_TC), against the CPU oracle
(VERDICT r1 "what's weak" #1, #2).

* cfgB: the bench scene itself, ``make_scene(3, 576, 800, 88, seed=1)`` (704^3 finest level: 349 M-entry index table,
  1.4 GB matching volume) — 16 chunks of 256 rays spread over the image (pixel row 0, the four corners) + one chunk of
  rays that miss the volume, rendered by both paths with the same jitter.
* cfgC: 8 random 64^3 blocks of the 512^3 extract_geometry grid.
* cfgD shape: 5 views (4 sources), 480x640, volumes 64 -> 512, 512 random-pixel rays.
* cfgE sampling: n_samples = [64, 32, 16, 16] (128 / ray), 5 views, 1080x1920 images, reduced-precision MLP mode at 1e-2.
* maximum sizes: an index table with more than 2^31 entries (1408^3, the Tanks-shaped finest level).

The oracle runs on the host cores: sizes are chosen so that every case finishes in seconds.
"""
import numpy as np
import pytest
import torch

import surf_oracle as O
from helpers import (RTOL_BF16, RTOL_FP32, assert_close, assert_equal_int, check_composited,
                     explain_gradient_mismatches, explain_mask_mismatches, explain_view_mismatches)
from surf_b200 import _lib, conf, synthetic
from surf_b200.modules import projector as P
from surf_b200.modules.implicit_surface import ImplicitSurface

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bench_net(mode=_lib.MLP_TC, n_samples=None, seed=0):
    """The network bench.py renders with (bench.build_net: geometric init + noise on the zero-initialised columns)."""
    import bench
    c = conf.default_implicit_surface_conf()
    if n_samples is not None:
        c.put("render.n_samples", list(n_samples))
    m = bench.build_net(None, seed=seed, confs=c)
    m.mlp_mode = mode
    return m.to(DEV)


def oracle_of(m, n_samples=(64, 32, 24, 16)):
    return O.OracleNet({k: v.detach().cpu() for k, v in m.state_dict().items()}, n_samples=n_samples)


def compare_chunk(m, net, ps, sc_cpu, o, d, t_rand, what, expect_hit=True, min_rows=0.85):
    """One reference render() call (a 256-ray chunk) by both paths with the same jitter; the full contract."""
    n = o.shape[0]
    near, far = sc_cpu.near.expand(n, 1), sc_cpu.far.expand(n, 1)
    pts_random = torch.rand(1024, 3, generator=torch.Generator().manual_seed(3)) * 2 - 1
    ref = O.render(net, o, d, near, far, sc_cpu.matching_volume, sc_cpu.volumes, sc_cpu.sparse_idxes,
                   sc_cpu.mask_volumes, sc_cpu.imgs, sc_cpu.features, sc_cpu.intrs, sc_cpu.c2ws, 1.0, t_rand=t_rand,
                   pts_random=pts_random, return_stages=True)
    out = m.render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV), ps, None, None, None, None, None, None, None, None,
                   1.0, None, t_rand=t_rand, pts_random=pts_random, return_stages=True)
    B, S = ref["mid_z_vals"].shape
    assert float((out["mid_z_vals"].cpu() - ref["mid_z_vals"]).abs().max()) <= 4e-6, what + "mid_z"
    rows = explain_mask_mismatches(O, out["_point_flags"].cpu() & 1, ref["_voxel_mask"], out["mid_z_vals"],
                                   ref["mid_z_vals"], o, d, sc_cpu.mask_volumes, what + "voxel mask")
    n_mask_mism = int((~rows).sum())
    # computed-point mask (= voxel mask, or the first 10 points of a chunk whose mask is empty, Q6): every ray without a
    # proven boundary mismatch must agree sample by sample — in an empty chunk that is every ray
    cm_gpu = ((out["_point_flags"].cpu() >> 1) & 1).bool().reshape(B, S)
    assert_equal_int(cm_gpu[rows], ref["_compute_mask"].reshape(B, S)[rows],
                     what + "computed-point mask (incl. the empty-chunk fallback)")
    rows = rows & explain_view_mismatches(O, out["_point_views"], ref, sc_cpu)
    g_rows, g_worst = explain_gradient_mismatches(O, out, ref, o, d, sc_cpu, net, what)
    rows = rows & g_rows
    same_cross = out["_prev_idx"].cpu().long() == ref["_prev_idx"][:, 0]
    rows = rows & same_cross
    assert float(rows.float().mean()) >= min_rows, what + "only %.2f of the rays are free of boundary cases" % float(rows.float().mean())
    assert_equal_int(out["valid_mask"].cpu()[rows], ref["valid_mask"][rows], what + "valid_mask")
    assert_equal_int(out["inside_sphere"].cpu()[rows], ref["inside_sphere"][rows], what + "inside_sphere")
    assert_equal_int(out["mid_inside_sphere"].cpu()[rows], ref["mid_inside_sphere"][rows], what + "mid_inside_sphere")
    assert_close(out["sparse_sdf"][:1024], ref["sparse_sdf"][:1024], RTOL_FP32, what + "sparse_sdf (random points)")
    worst = check_composited(O, out, ref, sc_cpu, net, o, d, rows=rows, what=what, grad_checked=True)
    worst["gradient"] = g_worst
    if expect_hit:
        assert float(ref["weight_sum"].max()) > 0.5, what + "the chunk should contain rays that hit the surface"
    return {"rays": B, "compared": int(rows.sum()), "mask_mismatches": n_mask_mism, "worst_err_over_tol": worst}


# ------------------------------------------------------------------------------------------------
# cfgB + cfgC: the bench scene
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def bench_scene():
    sc = synthetic.make_scene(3, 576, 800, 88, seed=1, device=DEV)
    m = bench_net()
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    sc_cpu = sc.to("cpu")
    rays_o, rays_d, hw = synthetic.image_rays(sc_cpu, 1)
    del sc
    torch.cuda.empty_cache()
    yield m, ps, sc_cpu, rays_o, rays_d
    ps.destroy()
    torch.cuda.empty_cache()


def test_bench_scene_chunks_vs_oracle(bench_scene):
    m, ps, sc_cpu, rays_o, rays_d = bench_scene
    net = oracle_of(m)
    n = rays_o.shape[0]
    assert n == 576 * 800
    n_chunks = n // 256
    # row 0 / top-left corner, top-right corner, bottom-left corner, last chunk = bottom-right corner, 12 spread
    chunks = sorted({0, 799 // 256, (575 * 800) // 256, n_chunks - 1} | {int(k * (n_chunks - 1) / 13.0) for k in range(1, 13)})
    assert len(chunks) >= 16
    torch.manual_seed(1234)
    t_all = m.draw_chunk_randoms(n)
    tot = {"rays": 0, "compared": 0, "mask_mismatches": 0}
    worst = {}
    for c in chunks:
        sl = slice(c * 256, (c + 1) * 256)
        # pixel row 0 projects to y = 0 +- rounding in both source views (they differ from the reference view by an x
        # offset only): the per-view validity bit 0 <= y (projector.py:536) is a coin flip in the reference itself, so
        # most rays of those chunks are census-excluded (with proof, explain_view_mismatches)
        r = compare_chunk(m, net, ps, sc_cpu, rays_o[sl], rays_d[sl], t_all[sl], "chunk %d: " % c, expect_hit=False,
                          min_rows=0.2 if c * 256 < 800 else 0.9)
        for k in tot:
            tot[k] += r[k]
        for k, v in r["worst_err_over_tol"].items():
            worst[k] = max(worst.get(k, 0.0), v)
    print("bench-scene parity: %s worst err/tol %s" % (tot, {k: round(v, 3) for k, v in worst.items()}))
    assert tot["compared"] >= 0.85 * tot["rays"]


def test_bench_scene_rays_that_miss(bench_scene):
    """A whole chunk of rays that never enter the volume: empty voxel mask -> the reference evaluates the first 10
    points anyway (implicit_surface.py:88-89) and composites nothing."""
    m, ps, sc_cpu, rays_o, rays_d = bench_scene
    net = oracle_of(m)
    g = torch.Generator().manual_seed(5)
    d = torch.nn.functional.normalize(torch.tensor([[1.0, 0.2, 0.1]]) + 0.05 * torch.randn(256, 3, generator=g), dim=1)
    o = rays_o[:256].clone()
    t = torch.rand(256, 4, generator=g)
    r = compare_chunk(m, net, ps, sc_cpu, o, d, t, "miss: ", expect_hit=False, min_rows=0.95)
    assert r["mask_mismatches"] == 0


def test_bench_scene_matches_bench_step(bench_scene):
    """The image pass bench.py times (render_image, 65 536-ray launch sets) equals the chunk-by-chunk reference calls
    bit for bit on the compared chunks."""
    m, ps, sc_cpu, rays_o, rays_d = bench_scene
    n = 4 * 256
    torch.manual_seed(1234)
    t_all = m.draw_chunk_randoms(rays_o.shape[0])
    near, far = sc_cpu.near.to(DEV), sc_cpu.far.to(DEV)
    img = m.render_image(ps, rays_o[:n].to(DEV), rays_d[:n].to(DEV), near, far, t_rand=t_all[:n])
    for c in range(4):
        sl = slice(c * 256, (c + 1) * 256)
        one = m.render(rays_o[sl].to(DEV), rays_d[sl].to(DEV), near.expand(256, 1), far.expand(256, 1), ps, None, None,
                       None, None, None, None, None, None, 1.0, None, t_rand=t_all[sl], pts_random=torch.zeros(1, 3))
        assert torch.equal(one["color_fine"], img["color_fine"][sl])
        assert torch.equal(one["render_depth"], img["render_depth"][sl])
        assert torch.equal(one["sdf_depth"], img["sdf_depth"][sl])


def test_grid_512_blocks_vs_oracle(bench_scene):
    """cfgC: the dense 512^3 SDF grid; 8 random 64^3 blocks (the reference's own block size) against the oracle."""
    m, ps, sc_cpu, _, _ = bench_scene
    net = oracle_of(m)
    u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 512)
    assert u.shape == (512, 512, 512)
    g = torch.Generator().manual_seed(7)
    blocks = torch.randint(0, 8, (8, 3), generator=g).tolist()
    blocks[0] = [3, 3, 3]            # a block that straddles the r = 0.5 surface and the finest shell
    worst = 0.0
    for bx, by, bz in blocks:
        xr, yr, zr = (bx * 64, bx * 64 + 64), (by * 64, by * 64 + 64), (bz * 64, bz * 64 + 64)
        ref = O.sdf_grid(net, sc_cpu.volumes, sc_cpu.sparse_idxes, [-1, -1, -1], [1, 1, 1], 512, xr, yr, zr)
        got = u[xr[0]:xr[1], yr[0]:yr[1], zr[0]:zr[1]].cpu()
        assert_close(got, ref, RTOL_FP32, "u block (%d,%d,%d)" % (bx, by, bz))
        worst = max(worst, float((got - ref).abs().max() / ref.abs().max()))
    print("512^3 grid blocks: worst relative error %.2e" % worst)
    # slab sharding (the multi-GPU partition) reproduces the grid bit for bit
    slab = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 512, x_range=(192, 256))
    assert torch.equal(slab, u[192:256])


# ------------------------------------------------------------------------------------------------
# cfgD shape: 5 views, 480x640, volumes 64 -> 512, 512 rays
# ------------------------------------------------------------------------------------------------
def test_cfgD_shape_vs_oracle():
    sc = synthetic.make_scene(5, 480, 640, 64, seed=10, device=DEV)
    m = bench_net()
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    sc_cpu = sc.to("cpu")
    del sc
    net = oracle_of(m)
    o, d = synthetic.random_pixel_rays(sc_cpu, 512, seed=12)
    t = torch.rand(512, 4, generator=torch.Generator().manual_seed(13))
    r = compare_chunk(m, net, ps, sc_cpu, o, d, t, "cfgD: ", min_rows=0.8)
    print("cfgD-shaped parity:", r)
    ps.destroy()


# ------------------------------------------------------------------------------------------------
# cfgE sampling: 128 samples / ray, 5 views, 1080x1920, reduced-precision MLP mode (1e-2)
# ------------------------------------------------------------------------------------------------
def test_cfgE_sampling_fast_mode_vs_oracle():
    ns = (64, 32, 16, 16)
    sc = synthetic.make_scene(5, 1080, 1920, 44, seed=20, device=DEV)
    m = bench_net(_lib.MLP_TC_FAST, n_samples=ns)
    assert sum(m.n_samples) == 128
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    sc_cpu = sc.to("cpu")
    del sc
    net = oracle_of(m, ns)
    o, d = synthetic.random_pixel_rays(sc_cpu, 256, seed=21)
    t = torch.rand(256, 4, generator=torch.Generator().manual_seed(22))
    near, far = sc_cpu.near.expand(256, 1), sc_cpu.far.expand(256, 1)
    pr = torch.zeros(1, 3)
    ref = O.render(net, o, d, near, far, sc_cpu.matching_volume, sc_cpu.volumes, sc_cpu.sparse_idxes, sc_cpu.mask_volumes,
                   sc_cpu.imgs, sc_cpu.features, sc_cpu.intrs, sc_cpu.c2ws, 1.0, t_rand=t, pts_random=pr, return_stages=True)
    out = m.render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV), ps, None, None, None, None, None, None, None, None,
                   1.0, None, t_rand=t, pts_random=pr, return_stages=True)
    assert out["mid_z_vals"].shape == (256, 128)
    rows = explain_mask_mismatches(O, out["_point_flags"].cpu() & 1, ref["_voxel_mask"], out["mid_z_vals"],
                                   ref["mid_z_vals"], o, d, sc_cpu.mask_volumes)
    cm = ref["_compute_mask"]
    keep_p = rows[:, None].expand(256, 128).reshape(-1) & cm
    assert_close(out["sparse_sdf"][1:].cpu()[keep_p], ref["_sdf"][keep_p], RTOL_BF16, "fast-mode per-point sdf")
    assert_close(out["gradients"].cpu().reshape(-1, 3)[keep_p], ref["_grad"][keep_p], RTOL_BF16, "fast-mode gradient")
    # inv_s = 20 at init: the composited outputs of the reduced-precision mode stay within 1e-2 as well
    for k in ("color_fine", "render_depth", "normal", "weight_sum"):
        assert_close(out[k].cpu()[rows], ref[k][rows], RTOL_BF16, "fast-mode " + k, floor=1e-2)
    ps.destroy()


# ------------------------------------------------------------------------------------------------
# maximum sizes: index table with more than 2^31 entries (Tanks-shaped finest level, 1408^3)
# ------------------------------------------------------------------------------------------------
def test_index_table_beyond_2_31_entries():
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~40 GB of device memory")
    N = 1408
    assert N ** 3 > 2 ** 31
    idx = torch.full((N, N, N), -1, dtype=torch.int64, device=DEV)
    z0 = 1200                                              # linear addresses z * N^2 + ... > 2^31 from z = 1084 on
    sub = idx[z0:z0 + 64, 600:700, 600:700]
    nvox = sub.numel()
    sub.copy_(torch.arange(nvox, device=DEV).reshape(sub.shape))
    g = torch.Generator(device=DEV).manual_seed(1)
    vol = torch.randn((nvox, 7), generator=g, device=DEV) * 0.1
    ps = P.PreparedScene([vol], [idx])
    del idx
    torch.cuda.empty_cache()
    # points inside that block: grid (x,y,z) <- world (z,y,x) (projector.py:379); table axes (D,H,W) = index[z][y][x]
    vs = 2.0 / (N - 1)
    n = 20000
    gz = (torch.rand(n, generator=g, device=DEV) * 62 + z0 + 0.5)
    gy = torch.rand(n, generator=g, device=DEV) * 98 + 600.5
    gx = torch.rand(n, generator=g, device=DEV) * 98 + 600.5
    pts = torch.stack([gz * vs - 1.0, gy * vs - 1.0, gx * vs - 1.0], dim=1)    # world x <-> table axis 0
    got = P.lookup_sparse_volume(pts, ps)
    # float64 restatement on the device (the CPU oracle would need the 22 GB table on the host)
    c = (pts.flip(-1).double() + 1.0) / (2.0 / (N - 1))
    c = ((pts.flip(-1) + 1.0) / torch.tensor(vs, device=DEV)).double()         # fp32 coordinates as the kernel forms them
    f0 = torch.floor(c)
    want = torch.zeros(n, 7, dtype=torch.float64, device=DEV)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                ix, iy, iz = f0[:, 0] + dx, f0[:, 1] + dy, f0[:, 2] + dz
                w = ((c[:, 0] - f0[:, 0]) if dx else (f0[:, 0] + 1 - c[:, 0])) * \
                    ((c[:, 1] - f0[:, 1]) if dy else (f0[:, 1] + 1 - c[:, 1])) * \
                    ((c[:, 2] - f0[:, 2]) if dz else (f0[:, 2] + 1 - c[:, 2]))
                # linear index iz * N^2 + iy * N + ix addresses table[iz][iy][ix]; rows were numbered in that block
                inb = (iz >= z0) & (iz < z0 + 64) & (iy >= 600) & (iy < 700) & (ix >= 600) & (ix < 700)
                row = ((iz - z0) * 100 + (iy - 600)) * 100 + (ix - 600)
                val = torch.where(inb[:, None], vol.double()[row.clamp(0, nvox - 1).long()], torch.zeros_like(want))
                want += val * w[:, None]
    assert float(want.abs().max()) > 0.05
    assert_close(got, want, 1e-5, "sparse gather above 2^31 linear addresses")
    ps.destroy()


# ------------------------------------------------------------------------------------------------
# other view counts: V = 3 and V = 1 source views run the block-synchronous blending kernel (k_blend_tc), V = 2 / 4 the
# TMEM-resident one (k_blend_tm); the goldens only hold V = 2 and V = 4
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nv", [4, 2])
def test_other_view_counts_vs_oracle(nv):
    sc = synthetic.make_scene(nv, 96, 128, 16, seed=30 + nv, device=DEV)
    m = bench_net()
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    sc_cpu = sc.to("cpu")
    del sc
    net = oracle_of(m)
    o, d = synthetic.random_pixel_rays(sc_cpu, 256, seed=5)
    t = torch.rand(256, 4, generator=torch.Generator().manual_seed(6))
    r = compare_chunk(m, net, ps, sc_cpu, o, d, t, "nv=%d: " % nv, min_rows=0.8)
    print("nv=%d parity:" % nv, r)
    ps.destroy()


# ------------------------------------------------------------------------------------------------
# the three schedules of the colour path (surf_render_cfg.color_path): serial (default), gather kernel beside the SDF
# kernel on a side stream, and gather fused into the blending kernel — same arithmetic per value, so bit-identical
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nv,mode", [(5, _lib.MLP_TC), (3, _lib.MLP_TC), (5, _lib.MLP_TC_FAST), (4, _lib.MLP_TC),
                                     (5, _lib.MLP_FFMA)])
def test_colour_path_schedules_are_bit_identical(nv, mode):
    sc = synthetic.make_scene(nv, 192, 256, 24, seed=50 + nv, device=DEV)
    m = bench_net()
    m.mlp_mode = mode
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    sc_cpu = sc.to("cpu")
    del sc
    n = 3000                                # ~400k points: many tiles per group, ragged last tile
    o, d = synthetic.random_pixel_rays(sc_cpu, n, seed=9)
    near, far = sc_cpu.near.expand(n, 1).to(DEV), sc_cpu.far.expand(n, 1).to(DEV)
    t = torch.rand(n, 4, generator=torch.Generator().manual_seed(8))
    pr = torch.rand(1024, 3, generator=torch.Generator().manual_seed(3)) * 2 - 1
    outs = []
    for path in (_lib.COLOR_SERIAL, _lib.COLOR_OVERLAP, _lib.COLOR_FUSED, _lib.COLOR_OVERLAP):
        m.color_path = path
        outs.append(m.render(o.to(DEV), d.to(DEV), near, far, ps, None, None, None, None, None, None, None, None, 1.0,
                             None, t_rand=t, pts_random=pr, return_stages=True))
    m.color_path = _lib.COLOR_SERIAL
    a = outs[0]
    assert int(a["_point_views"].ne(0).sum()) > 10000, "the scene should project into the source views"
    for i, b in enumerate(outs[1:]):
        for k in ("_point_color", "_point_views", "color_fine", "weights", "gradients", "sdf_depth"):
            assert torch.equal(a[k], b[k]), "colour path %d differs from serial in %s (V=%d, mode %d)" % (i, k, nv - 1, mode)
    ps.destroy()
