"""Parity of the CUDA path (through the C-ABI, via the module mirror) against the CPU oracle and the
reference-generated golden vectors.

Bars (BASELINE.json north_star): bit-exact for sample indices / sparsity masks in stage-isolated mode
(oracle floats in); fp32 outputs within 1e-4 relative (scale-relative, helpers.assert_close).
End-to-end integer outputs use a mismatch census: reductions (softmax / 4x4 inverses / K=4 matmuls)
cannot be made order-identical to MKL, so a sample within an ulp of a voxel or pixel boundary may flip;
every mismatch must be explained by such a boundary and the count must stay tiny.
"""
import numpy as np
import pytest
import torch

import surf_oracle as O
from helpers import (RTOL_FP32, assert_close, assert_equal_int, blend_envelope, check_composited, explain_mask_mismatches,
                     explain_view_mismatches, load_golden, scene_from_recipe)
from surf_b200 import _lib, conf, synthetic
from surf_b200.modules import projector as P
from surf_b200.modules.implicit_surface import ImplicitSurface

pytestmark = pytest.mark.gpu

RENDER_CASES = ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed", "render_miss"]
DEV = "cuda:0"


# every parity test runs in both fp32-grade kernel families: the tcgen05 kernels (the default of the drop-in module and
# what bench.py measures) and the fp32 FFMA parity anchor
MODES = [_lib.MLP_TC, _lib.MLP_FFMA]


@pytest.fixture(params=MODES, ids=["tc", "ffma"])
def mode(request):
    return request.param


def build(g, mode=_lib.MLP_TC):
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    assert m.mlp_mode == _lib.MLP_TC, "the drop-in module must default to the benchmarked tcgen05 kernels"
    m.load_state_dict(g["sd"], strict=True)
    m.mlp_mode = mode
    return m.to(DEV)


def gpu_scene(m, sc):
    d = sc.to(DEV)
    return d, m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs,
                        d.c2ws)


def oracle_render(g, sc, stages=True, t_rand=None, pts_random=None):
    net = O.OracleNet(g["sd"])
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    return net, O.render(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, sc.volumes,
                         sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0,
                         t_rand=t_rand, pts_random=pts_random, return_stages=stages)


# ------------------------------------------------------------------------------------------------
# stage-isolated: masks and gathers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", RENDER_CASES[:3])
def test_voxel_mask_bit_exact(name):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = build(g)
    _, ps = gpu_scene(m, sc)
    _, ref = oracle_render(g, sc)
    gen = torch.Generator().manual_seed(5)
    extra = torch.rand(200000, 3, generator=gen) * 2.4 - 1.2          # incl. points outside the cube
    # points exactly on voxel boundaries of the finest level (round-half-even cases)
    N = sc.mask_volumes[0].shape[-1]
    edges = (torch.randint(0, 2 * N + 1, (20000, 3), generator=gen).float() / N) - 1.0
    pts = torch.cat([ref["_pts"], extra, edges])
    want = O.point_mask(pts, sc.mask_volumes)
    got = P.lookup_volume(pts.to(DEV), ps)[:, 0] > 0
    assert_equal_int(got, want, "voxel mask")
    assert 0.02 < float(want.float().mean()) < 0.98


@pytest.mark.parametrize("name", RENDER_CASES[:3])
def test_sparse_gather(name):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = build(g)
    _, ps = gpu_scene(m, sc)
    pv = torch.from_numpy(g["out"]["_pts_valid"])
    got = P.lookup_sparse_volume(pv.to(DEV), ps)
    assert_close(got, g["out"]["_sparse_feats"], 1e-5, "sparse feats vs reference golden")
    gen = torch.Generator().manual_seed(6)
    wild = torch.rand(50000, 3, generator=gen) * 3.0 - 1.5
    assert_close(P.lookup_sparse_volume(wild.to(DEV), ps), O.lookup_sparse(wild, sc.volumes, sc.sparse_idxes), 1e-5,
                 "sparse feats, extrapolating points")


@pytest.mark.parametrize("name", RENDER_CASES[:3])
def test_lookup_feature(name):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = build(g)
    _, ps = gpu_scene(m, sc)
    pv = torch.from_numpy(g["out"]["_pts_valid"])
    fv, rd, mk = P.lookup_feature(pv.to(DEV), ps)
    ref_mask = torch.from_numpy(g["out"]["_view_mask"])
    mism = (mk.cpu() != ref_mask)
    # census: the projection uses MKL batched matmuls in the reference; a mismatch is only legal for
    # a projection within 1e-3 px of an image border
    assert int(mism.sum()) <= max(2, int(1e-4 * mism.numel())), "view-mask mismatches: %d" % int(mism.sum())
    assert_close(rd, g["out"]["_ray_diff"], 1e-5, "ray_diff")
    ok = ~mism.any(dim=1)
    assert_close(fv.cpu()[ok], torch.from_numpy(g["out"]["_feat_views"])[ok], RTOL_FP32, "feat_views")


@pytest.mark.parametrize("name", RENDER_CASES[:3])
def test_blend(name, mode):
    g = load_golden(name)
    m = build(g, mode)
    o = g["out"]
    net = O.OracleNet(g["sd"])
    fv, rd, mk = (torch.from_numpy(o[k]) for k in ("_feat_views", "_ray_diff", "_view_mask"))
    got = m.color_network(fv.to(DEV), rd.to(DEV), mk.to(DEV)).cpu()
    ref = torch.from_numpy(o["_blend_rgb"])
    # 1e-4 relative, widened only where the reference itself is ill-conditioned (helpers.blend_envelope)
    _, env = blend_envelope(O, net, fv, rd, mk)
    err = (got - ref).abs().max(dim=1)[0]
    tol = RTOL_FP32 * float(ref.abs().max()) + 2.0 * env
    assert bool((err <= tol).all()), "blend rgb: %d/%d points beyond 1e-4 + envelope (max err %.3e)" % (
        int((err > tol).sum()), err.numel(), float(err.max()))
    well = env < 1e-6          # with 4 near-symmetric source views only a minority of points is well-conditioned
    assert float(well.float().mean()) > (0.5 if fv.shape[1] <= 2 else 0.05)
    assert_close(got[well], ref[well], RTOL_FP32, "blend rgb, well-conditioned points")
    # all views masked -> uniform softmax over the raw samples
    mk0 = torch.zeros_like(mk)
    want = O.blend(net, fv, rd, mk0)
    assert_close(m.color_network(fv.to(DEV), rd.to(DEV), mk0.to(DEV)), want, RTOL_FP32, "blend rgb, nothing visible")
    assert torch.equal(got, m.color_network(fv.to(DEV), rd.to(DEV), mk.to(DEV)).cpu()), "blend must be deterministic"


# ------------------------------------------------------------------------------------------------
# SDF MLP
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", RENDER_CASES[:3])
def test_sdf_and_gradient(name, mode):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = build(g, mode)
    _, ps = gpu_scene(m, sc)
    pv = torch.from_numpy(g["out"]["_pts_valid"]).to(DEV)
    sdf = m.sdf_network.sdf(pv, ps)
    assert_close(sdf, g["out"]["_sdf_full"][:, :1], RTOL_FP32, "sdf vs reference golden")
    s2, grad = m.sdf_network.gradient(pv, ps, with_sdf=True)
    assert_close(s2, g["out"]["_sdf_full"][:, :1], RTOL_FP32, "sdf (gradient kernel)")
    assert_close(grad, g["out"]["_grad_valid"], RTOL_FP32, "d sdf / d x vs reference autograd")
    assert torch.equal(sdf, s2), "forward-only and forward+reverse kernels must agree bit-for-bit"
    s3, g3 = m.sdf_network.gradient(pv, ps, with_sdf=True)
    assert torch.equal(s2, s3) and torch.equal(grad, g3), "the gradient kernel must be deterministic"
    for n in (1, 127, 129, 300):
        s4, g4 = m.sdf_network.gradient(pv[:n], ps, with_sdf=True)
        assert torch.equal(s4, s2[:n]) and torch.equal(g4, grad[:n]), "results must not depend on the tile packing"


def test_sdf_modes_agree():
    """The two fp32-grade kernel families agree far inside the tolerance (measured 3e-6 of scale)."""
    g = load_golden("render_v2_perturbed")
    sc = scene_from_recipe(g["recipe"])
    m = build(g)
    _, ps = gpu_scene(m, sc)
    pv = torch.from_numpy(g["out"]["_pts_valid"]).to(DEV)
    m.mlp_mode = _lib.MLP_TC
    a, ga = m.sdf_network.gradient(pv, ps, with_sdf=True)
    m.mlp_mode = _lib.MLP_FFMA
    b, gb = m.sdf_network.gradient(pv, ps, with_sdf=True)
    assert float((a - b).abs().max()) < 2e-5 * float(b.abs().max())
    assert float((ga - gb).abs().max()) < 5e-5 * float(gb.abs().max())
    assert not torch.equal(a, b), "the mode switch must select different kernels"


def test_sdf_wild_points_and_ragged_sizes(mode):
    g = load_golden("sdf_grid_24")
    sc = scene_from_recipe(g["recipe"])
    m = build(g, mode)
    _, ps = gpu_scene(m, sc)
    wild = g["in"]["wild_pts"].to(DEV)
    s, gr = m.sdf_network.gradient(wild, ps, with_sdf=True)
    assert_close(s, g["out"]["wild_full"][:, :1], RTOL_FP32, "sdf, out-of-range points")
    assert_close(gr, g["out"]["wild_grad"], RTOL_FP32, "gradient, out-of-range points")
    # ragged / tiny / empty inputs
    reps = -(-1000 // wild.shape[0])
    wild_r, s_r = wild.repeat(reps, 1), s.repeat(reps, 1)
    for n in (0, 1, 127, 128, 129, 255, 257, 323, 1000):
        p = wild_r[:n]
        out = m.sdf_network.sdf(p, ps)
        assert out.shape == (n, 1)
        if n:
            assert torch.equal(out, s_r[:n])


def test_sdf_grid_matches_reference_and_slabs(mode):
    g = load_golden("sdf_grid_24")
    sc = scene_from_recipe(g["recipe"])
    m = build(g, mode)
    _, ps = gpu_scene(m, sc)
    res = int(g["recipe"]["resolution"])
    u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], res)
    assert_close(u, g["out"]["u"], RTOL_FP32, "u grid vs reference golden")
    slabs = torch.cat([m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], res, x_range=(a, b)) for a, b in ((0, 7), (7, 8), (8, 24))])
    assert torch.equal(slabs, u), "x-slab sharding must reproduce the full grid bit-for-bit"
    # opt-in sparsified mode: identical inside the mask, constant outside (Q16)
    us = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], res, sparsify=True, fill=-100.0)
    lin = torch.linspace(-1, 1, res)
    xx, yy, zz = torch.meshgrid(lin, lin, lin, indexing="ij")
    inside = O.point_mask(torch.stack([xx, yy, zz], -1).reshape(-1, 3), sc.mask_volumes).reshape(res, res, res)
    assert torch.equal(us.cpu()[inside], u.cpu()[inside])
    assert bool((us.cpu()[~inside] == -100.0).all())


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", RENDER_CASES)
def test_sample_z(name):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = build(g)
    d, ps = gpu_scene(m, sc)
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    t_rand = O.draw_t_rand(i["rays_o"].shape[0], 4)
    net = O.OracleNet(g["sd"])
    z_ref, surf_ref = O.sample_z(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, t_rand)
    z, surf = m.sample_z(ps, i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), t_rand)
    z = z.cpu()
    assert bool((z[:, 1:] >= z[:, :-1]).all()), "z_vals must be sorted"
    assert_close(surf, surf_ref[:, 0], 1e-5, "expected surface depth")
    assert float((z - z_ref).abs().max()) <= 4e-6, "z_vals differ by more than a few ulp"
    # stage 0 does not depend on the probe: those 64 values must be present bit-exactly
    lin = torch.linspace(0.0, 1.0, 64)
    z0 = i["near"] + (i["far"] - i["near"]) * lin[None] + (t_rand[:, 0:1] - 0.5) * 2.0 / 64
    for r in range(z.shape[0]):
        assert np.isin(z0[r].numpy(), z[r].numpy()).all()
    # no jitter path
    z_nj, _ = m.sample_z(ps, i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), None)
    z_nj_ref, _ = O.sample_z(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, None)
    assert float((z_nj.cpu() - z_nj_ref).abs().max()) <= 4e-6


# ------------------------------------------------------------------------------------------------
# render_core, stage-isolated: the oracle's z_vals go in, integers must be bit-exact
# ------------------------------------------------------------------------------------------------
INT_KEYS = ["valid_mask", "inside_sphere", "mid_inside_sphere"]
POINT_KEYS = ["gradients", "mid_z_vals", "s_val"]          # per-point values: plain 1e-4 (sdf: check_composited)


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_core_stage_isolated(name, mode):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = build(g, mode)
    d, ps = gpu_scene(m, sc)
    net, ref0 = oracle_render(g, sc)
    z_vals = ref0["_z_vals"]                       # the oracle's own sorted z_vals go into both paths
    i = g["in"]
    pts_random = torch.rand(1024, 3, generator=torch.Generator().manual_seed(9)) * 2 - 1
    ref = O.render_core(net, i["rays_o"], i["rays_d"], z_vals, sc.volumes, sc.sparse_idxes, sc.mask_volumes,
                        sc.features, sc.imgs, sc.intrs, sc.c2ws, 1.0, pts_random=pts_random, return_stages=True)
    out = m.render_core(i["rays_o"].to(DEV), i["rays_d"].to(DEV), z_vals.to(DEV), 2.0 / 64, ps,
                        None, None, None, None, None, None, None, None, None, 1.0, None, pts_random=pts_random,
                        return_stages=True)
    B, S = z_vals.shape
    flags = out["_point_flags"].cpu()
    assert_equal_int(flags & 1, ref["_voxel_mask"], "voxel mask (stage-isolated)")
    assert_equal_int((flags >> 1) & 1, ref["_compute_mask"], "computed-point mask incl. empty-chunk fallback")
    assert np.array_equal(out["mid_z_vals"].cpu().numpy(), ref["mid_z_vals"].numpy()), "mid_z must be bit-identical"
    for k in ("inside_sphere",):
        assert_equal_int(out[k], ref[k], k)
    # per-view validity: a mismatch is legal only for a projection within 1e-3 px of an image border
    rows = explain_view_mismatches(O, out["_point_views"], ref, sc)
    assert int(rows.sum()) >= B - 2
    assert_equal_int(out["valid_mask"].cpu()[rows], ref["valid_mask"][rows], "valid_mask")
    # first zero crossing: bit-exact unless the sign of an SDF value is itself inside the tolerance
    pi, pr = out["_prev_idx"].cpu().long(), ref["_prev_idx"][:, 0]
    diff = pi != pr
    if bool(diff.any()):
        sd = ref["_sdf"].reshape(B, S)
        lo = torch.minimum(pi, pr)[diff]
        near0 = torch.stack([sd[diff].gather(1, (lo[:, None] + k).clamp(max=S - 1)).abs()[:, 0] for k in range(3)], 1).min(1)[0]
        assert bool((near0 < RTOL_FP32 * float(sd[ref["_voxel_mask"].reshape(B, S)].abs().max())).all()), \
            "zero-crossing index differs although no SDF value is within tolerance of 0"
    rows = rows & ~diff
    assert int(rows.sum()) >= B - 3
    assert_equal_int(out["mid_inside_sphere"].cpu()[rows], ref["mid_inside_sphere"][rows], "mid_inside_sphere")
    for k in POINT_KEYS:
        assert_close(out[k], ref[k], RTOL_FP32, k)
    # sparse_sdf = [1024 random points (0 outside the mask) | per-sample sdf]
    assert_close(out["sparse_sdf"][:1024], ref["sparse_sdf"][:1024], RTOL_FP32, "sparse_sdf (random points)")
    check_composited(O, out, ref, sc, net, i["rays_o"], i["rays_d"], rows=rows, per_sample=True)
    ge_ref = float(ref["gradient_error"])
    assert abs(float(out["gradient_error"]) - ge_ref) <= 2e-4 * max(ge_ref, 1e-2)


# ------------------------------------------------------------------------------------------------
# end to end against the reference golden
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_end_to_end_vs_reference(name, mode):
    """render() with the reference's own RNG stream against the UNMODIFIED reference's outputs (golden)."""
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = build(g, mode)
    d, ps = gpu_scene(m, sc)
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))          # same global-RNG stream as the reference run
    out = m.render(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), d.matching_volume,
                   d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.features, d.intrs, d.c2ws, 1.0,
                   None, return_stages=True)
    # the oracle (pinned to the same golden by tests/test_oracle_golden.py) supplies the per-point stages
    net, ref = oracle_render(g, sc)
    gold = g["out"]
    for k in ("color_fine", "render_depth", "sdf_depth", "normal"):
        assert_close(ref[k], gold[k], 1e-5, "oracle vs reference golden: " + k, floor=1e-2)
    B, S = ref["mid_z_vals"].shape
    assert float((out["mid_z_vals"].cpu() - ref["mid_z_vals"]).abs().max()) <= 4e-6
    rows = explain_mask_mismatches(O, out["_point_flags"].cpu() & 1, torch.from_numpy(gold["_voxel_mask"]),
                                   out["mid_z_vals"], ref["mid_z_vals"], i["rays_o"], i["rays_d"], sc.mask_volumes)
    assert int(rows.sum()) >= B - 2
    rows = rows & explain_view_mismatches(O, out["_point_views"], ref, sc)
    # a sample on a voxel FACE has a discontinuous trilinear gradient (oracle.voxel_face_distance): the two paths'
    # sample positions differ by an ulp, so such rays are compared for everything but the gradient-driven outputs
    on_face = (O.voxel_face_distance(ref["_pts"], sc.sparse_idxes) < 1.5).reshape(B, S).any(dim=1)
    rows = rows & ~on_face
    assert int(rows.sum()) >= int(0.9 * B)
    assert_equal_int(out["valid_mask"].cpu()[rows], ref["valid_mask"][rows], "valid_mask")
    assert_equal_int(out["inside_sphere"].cpu()[rows], ref["inside_sphere"][rows], "inside_sphere")
    same_cross = out["_prev_idx"].cpu().long() == ref["_prev_idx"][:, 0]
    assert float(same_cross.float().mean()) > 0.95
    rows = rows & same_cross
    assert_equal_int(out["mid_inside_sphere"].cpu()[rows], ref["mid_inside_sphere"][rows], "mid_inside_sphere")
    check_composited(O, out, ref, sc, net, i["rays_o"], i["rays_d"], rows=rows)
    assert set(out.keys()) >= {"color_fine", "render_depth", "sdf_depth", "normal", "valid_mask", "sparse_sdf",
                               "mid_z_vals", "gradients", "s_val", "weights", "weight_sum", "weight_max",
                               "gradient_error", "inside_sphere", "mid_inside_sphere"}


def test_validate_image_vs_reference(mode):
    g = load_golden("validate_24x32")
    sc = scene_from_recipe(g["recipe"])
    m = build(g, mode)
    d, ps = gpu_scene(m, sc)
    i = g["in"]
    r = int(g["recipe"]["res_level"])
    hw = (sc.H // r, sc.W // r)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = m.validate(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), d.matching_volume,
                     d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.features, d.intrs, d.c2ws,
                     torch.tensor([-1.0, -1, -1]), torch.tensor([1.0, 1, 1]), hw, 1.0, None, extract_geometry=False)
    assert isinstance(out["color_fine"], torch.Tensor) and out["color_fine"].device.type == "cpu"
    assert isinstance(out["img_fine"], np.ndarray) and out["img_fine"].shape == (hw[0], hw[1], 3)
    n = hw[0] * hw[1]
    # the same image through the oracle, chunk by chunk with the reference's RNG stream, for the per-point stages
    net = O.OracleNet(g["sd"])
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    t_rand = m.draw_chunk_randoms(n)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    t2 = m.draw_chunk_randoms(n)
    assert torch.equal(t_rand, t2)
    res = m.render_image(ps, i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), t_rand=t_rand)
    for k in ("color_fine", "sdf_depth", "render_depth"):
        assert np.array_equal(res[k].cpu().numpy().reshape(-1), np.asarray(out[k]).reshape(-1)), \
            "validate() must be render_image() with the reference's chunked RNG stream (%s)" % k
    near, far = i["near"].expand(n, 1), i["far"].expand(n, 1)
    bad_rays = torch.zeros(n, dtype=torch.bool)
    worst = {}
    for c0 in range(0, n, 256):
        sl = slice(c0, min(n, c0 + 256))
        ref = O.render(net, i["rays_o"][sl], i["rays_d"][sl], near[sl], far[sl], sc.matching_volume, sc.volumes,
                       sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0,
                       t_rand=t_rand[sl], pts_random=torch.zeros(1, 3), return_stages=True)
        o2 = m.render(i["rays_o"][sl].to(DEV), i["rays_d"][sl].to(DEV), near[sl].to(DEV), far[sl].to(DEV), ps, None,
                      None, None, None, None, None, None, None, 1.0, None, t_rand=t_rand[sl],
                      pts_random=torch.zeros(1, 3), return_stages=True)
        B, S = ref["mid_z_vals"].shape
        rows = explain_mask_mismatches(O, o2["_point_flags"].cpu() & 1, ref["_voxel_mask"], o2["mid_z_vals"],
                                       ref["mid_z_vals"], i["rays_o"][sl], i["rays_d"][sl], sc.mask_volumes)
        rows = rows & explain_view_mismatches(O, o2["_point_views"], ref, sc)
        on_face = (O.voxel_face_distance(ref["_pts"], sc.sparse_idxes) < 1.5).reshape(B, S).any(dim=1)
        rows = rows & ~on_face & (o2["_prev_idx"].cpu().long() == ref["_prev_idx"][:, 0])
        bad_rays[sl] = ~rows
        w = check_composited(O, o2, ref, sc, net, i["rays_o"][sl], i["rays_d"][sl], rows=rows, what="chunk %d: " % c0)
        for k, v in w.items():
            worst[k] = max(worst.get(k, 0.0), v)
        # the chunked device pass equals the one-launch image pass bit for bit (per-chunk RNG and fallback kept)
        assert torch.equal(o2["color_fine"], res["color_fine"][sl])
    assert float(bad_rays.float().mean()) < 0.2, "too many rays excluded by boundary cases"
    # and the golden image of the UNMODIFIED reference: every ray not excluded above agrees within the loosest
    # composited tolerance observed (the reference golden has no per-point stages to derive an envelope from)
    for k in ["color_fine", "img_fine", "normal_img", "sdf_depth", "render_depth"]:
        a = torch.as_tensor(np.asarray(out[k])).double().reshape(n, -1)
        b = torch.as_tensor(np.asarray(g["out"][k])).double().reshape(n, -1)
        scale = max(float(b.abs().max()), 1e-2)
        err = ((a - b).abs() / scale).max(dim=1)[0]
        assert float(err[~bad_rays].max()) < 2e-3, "%s: %.3e of scale vs the reference image" % (k, float(err[~bad_rays].max()))


# ------------------------------------------------------------------------------------------------
# size-independent properties on a larger scene
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mlp_mode", [_lib.MLP_FFMA, _lib.MLP_TC, _lib.MLP_TC_FAST])
def test_properties_larger_scene(mlp_mode):
    """Size-independent properties in every kernel family."""
    sc = synthetic.make_scene(3, 144, 200, 16, seed=4)
    g = load_golden("render_v2_perturbed")
    m = build(g, mlp_mode)
    d, ps = gpu_scene(m, sc)
    o, dd, hw = synthetic.image_rays(sc, 2)            # 72 x 100 = 7200 rays, 29 chunks (last one ragged)
    o, dd = o.to(DEV), dd.to(DEV)
    near, far = d.near, d.far
    torch.manual_seed(11)
    t_rand = m.draw_chunk_randoms(o.shape[0])
    full = m.render_image(ps, o, dd, near, far, t_rand=t_rand)
    again = m.render_image(ps, o, dd, near, far, t_rand=t_rand)
    for k in full:
        assert torch.equal(full[k], again[k]), "render must be deterministic (%s)" % k
    # ray sharding on chunk boundaries == whole image, bit for bit (the multi-GPU contract)
    cut = 256 * 13
    m.ray_batch = 256 * 4
    a = m.render_image(ps, o[:cut], dd[:cut], near, far, t_rand=t_rand[:cut])
    b = m.render_image(ps, o[cut:], dd[cut:], near, far, t_rand=t_rand[cut:])
    for k in full:
        assert torch.equal(torch.cat([a[k], b[k]]), full[k]), "sharded render differs (%s)" % k
    assert bool(torch.isfinite(full["color_fine"]).all())
    c = full["color_fine"]
    assert float(c.min()) >= -1e-5 and float(c.max()) <= 1.0 + 1e-4      # convex blend of images in [0,1)
    assert float((full["sdf_depth"] > 0).float().mean()) > 0.05            # rays do hit the r=0.5 sphere
    empty = m.render_image(ps, o[:0], dd[:0], near, far, t_rand=t_rand[:0])
    assert empty["color_fine"].shape == (0, 3) and empty["render_depth"].shape == (0,)


def test_scene_cache_keeps_live_scenes():
    """stage helpers and cache eviction must never destroy a scene a caller still holds, and new views
    must not re-prepare the volume part."""
    from surf_b200.scene import GLOBAL_SCENE_CACHE
    g = load_golden("render_v2_perturbed")
    sc = scene_from_recipe(g["recipe"])
    m = build(g)
    d, ps = gpu_scene(m, sc)
    pts = torch.rand(64, 3, device=DEV) * 2 - 1
    for _ in range(8):      # raw-tensor stage helpers build throw-away scenes
        P.lookup_volume(pts, [mv.clone() for mv in d.mask_volumes])
        P.lookup_feature(pts, d.imgs.clone(), d.intrs, d.c2ws, [f.clone() for f in d.features])
    others = [synthetic.make_scene(3, 48, 64, 8, seed=50 + k).to(DEV) for k in range(6)]
    for s2 in others:       # more scenes than the LRU holds
        m.prepare(s2.matching_volume, s2.volumes, s2.sparse_idxes, s2.mask_volumes, s2.imgs, s2.features, s2.intrs, s2.c2ws)
    assert ps._h is not None
    i = g["in"]
    out = m.render(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), ps, None, None, None,
                   None, None, None, None, None, 1.0, None)
    assert bool(torch.isfinite(out["color_fine"]).all())
    # new views on the same volumes: same scene object, volume part untouched
    GLOBAL_SCENE_CACHE.clear()
    a = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    feats2 = [f.clone() for f in d.features]
    b = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, feats2, d.intrs, d.c2ws)
    assert a is b
