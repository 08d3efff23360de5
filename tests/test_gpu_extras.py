"""Training extras of render() on the GPU (csrc/extras.cu): surface point, its normal (third MLP pass) and
surface_patch_warp2 (projector.py:560-645; implicit_surface.py:218-245), stage by stage against the oracle and end to end
against the reference's own outputs (goldens)."""
import numpy as np
import pytest
import torch

import surf_oracle as O
from helpers import RTOL_FP32, assert_close, load_golden, scene_from_recipe
from surf_b200 import _lib, conf
from surf_b200.modules.implicit_surface import ImplicitSurface

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TRAIN_KEYS = {"ref_gray_val", "sampled_gray_val", "mid_inside_sphere", "color_fine", "render_depth", "valid_mask",
              "sparse_sdf", "mid_z_vals", "gradients", "normal", "s_val", "weights", "weight_sum", "weight_max",
              "gradient_error", "inside_sphere", "sdf_depth", "smooth_error"}        # the reference's 18 keys


def _render(name, mode):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.mlp_mode = mode
    m.load_state_dict(g["sd"])
    m = m.to(DEV)
    d = sc.to(DEV)
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = m.render(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), d.matching_volume,
                   d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.features, d.intrs, d.c2ws, 1.0, None)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    net = O.OracleNet(g["sd"])
    ref = O.render(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, sc.volumes, sc.sparse_idxes,
                   sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0, extras=True, return_stages=True)
    return g, sc, net, out, ref


@pytest.mark.parametrize("mode", [_lib.MLP_TC, _lib.MLP_FFMA], ids=["tc", "ffma"])
@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed", "render_miss"])
def test_patch_warp_stage_by_stage(name, mode):
    g, sc, net, out, ref = _render(name, mode)
    assert TRAIN_KEYS <= set(out.keys())
    B = ref["mid_z_vals"].shape[0]
    V = sc.nv - 1
    assert out["ref_gray_val"].shape == (1, B, 121, 12) and out["sampled_gray_val"].shape == (V, B, 121, 12)
    valid = (ref["mid_inside_sphere"].reshape(-1) > 0) & (out["mid_inside_sphere"].cpu().reshape(-1) > 0)
    # (1) the surface point: a linear interpolation of two samples (1e-4 of the scene scale for rays with a crossing)
    p_gpu = out["_pts_sdf0"].cpu()
    if bool(valid.any()):
        assert float((p_gpu[valid] - ref["_pts_sdf0"][valid]).abs().max()) <= 1e-4 * 3.0, "pts_sdf0"
    # (2) the normal, stage-isolated: the oracle's gradient AT THE GPU's surface points, normalised and rotated
    _, g_o = O.sdf_gradient(net, p_gpu, sc.volumes, sc.sparse_idxes)
    gn = torch.linalg.norm(g_o, ord=2, dim=-1, keepdim=True)
    n_o = torch.matmul(sc.c2ws[0, :3, :3].permute(1, 0)[None], (g_o / gn.clamp(min=1e-8))[:, :, None]).squeeze(-1)
    scale = float(g_o.abs().max())
    tol = 2.0 * RTOL_FP32 * scale / gn.clamp(min=1e-8)            # gradient at 1e-4 of scale -> unit vector
    err_n = (out["_normal_sdf0"].cpu() - n_o).abs().max(dim=1, keepdim=True)[0]
    well = (gn > 1e-2 * scale).reshape(-1)
    assert bool((err_n[well] <= tol[well]).all()), "normal at the surface point: max err %.3e" % float(err_n[well].max())
    # (3) the patch warp itself, stage-isolated: oracle surface_patch_warp2 on the GPU's points and normals;
    #     tolerance 1e-4 of scale + the change of the oracle's own samples when a pixel coordinate moves by 4 ulp
    wf = O.warp_feature_maps(sc.features)
    pts3, nrm3 = p_gpu.reshape(B, 1, 3), out["_normal_sdf0"].cpu().reshape(B, 1, 3)
    r0, s0, dbg = O.surface_patch_warp2(pts3, nrm3, wf, sc.intrs, sc.c2ws)
    ulp = 4 * np.finfo(np.float32).eps * max(sc.H, sc.W)
    env_r, env_s = torch.zeros_like(r0), torch.zeros_like(s0)
    for sh in ((ulp, 0.0), (0.0, ulp), (-ulp, -ulp)):
        r1, s1, _ = O.surface_patch_warp2(pts3, nrm3, wf, sc.intrs, sc.c2ws, pixel_shift=sh)
        env_r, env_s = torch.maximum(env_r, (r1 - r0).abs()), torch.maximum(env_s, (s1 - s0).abs())
    sc_r, sc_s = float(r0.abs().max()), float(s0[torch.isfinite(s0)].abs().max())
    er = (out["ref_gray_val"].cpu() - r0).abs()
    assert bool((er <= RTOL_FP32 * sc_r + 2 * env_r).all()), "ref_gray_val: max err %.3e (scale %.2f)" % (float(er.max()), sc_r)
    # source views: rays whose plane-induced homography is well conditioned (|n . p| not tiny) and finite in the oracle
    cond = (dbg["disp"].abs() > 1e-3) & torch.isfinite(s0).reshape(V, B, -1).all(dim=2).all(dim=0)
    es = (out["sampled_gray_val"].cpu() - s0).abs()[:, cond]
    assert bool((es <= RTOL_FP32 * sc_s + 2 * env_s[:, cond]).all()), "sampled_gray_val: max err %.3e (scale %.2f)" % (
        float(es.max()), sc_s)
    assert int((cond & valid).sum()) >= 0.9 * int(valid.sum()), "the well-conditioned set must cover the rays that matter"
    # (4) end to end against the reference's own outputs, rays with a valid crossing.  The chain point -> normal ->
    #     homography amplifies the admitted deviations of (1) and (2) (the softplus(100 x) network is almost piecewise
    #     linear: its gradient jumps across kinks a few 1e-5 apart; a normal error dn moves a warped pixel by
    #     ~ f |baseline| / depth * dn and the feature maps change by O(1) per pixel), so the tolerance is 1e-4 of scale
    #     + 1.5 x the change of the ORACLE's warp between the reference's (point, normal) and the GPU's
    if bool(valid.any()):
        gr, gs = torch.as_tensor(g["out"]["ref_gray_val"]), torch.as_tensor(g["out"]["sampled_gray_val"])
        r_ref, s_ref, _ = O.surface_patch_warp2(ref["_pts_sdf0"].reshape(B, 1, 3), ref["_normal_sdf0"].reshape(B, 1, 3), wf,
                                                sc.intrs, sc.c2ws)
        assert torch.equal(r_ref, gr) and torch.equal(s_ref[:, valid], gs[:, valid]), "oracle == reference golden"
        vc = valid & cond
        e_r = (out["ref_gray_val"].cpu() - gr).abs()[:, valid]
        t_r = (RTOL_FP32 * sc_r + 2 * env_r + 1.5 * (r0 - r_ref).abs())[:, valid]
        assert bool((e_r <= t_r).all()), "ref_gray_val vs reference: max err %.3e" % float(e_r.max())
        e_s = (out["sampled_gray_val"].cpu() - gs).abs()[:, vc]
        t_s = (RTOL_FP32 * sc_s + 2 * env_s + 1.5 * (s0 - s_ref).abs())[:, vc]
        assert bool((e_s <= t_s).all()), "sampled_gray_val vs reference: max err %.3e" % float(e_s.max())
        print("%s mode %d: %d valid rays; sampled_gray_val max |gpu - reference| %.2e (scale %.2f), of which the kernel "
              "itself %.2e" % (name, mode, int(vc.sum()), float(e_s.max()), sc_s, float(es.max())))


def test_warp_maps_follow_the_views():
    """The 12-channel maps live behind the scene handle and are rebuilt when surf_scene_set_views installs other views
    (SuRF.forward with view_ids): same result as a freshly prepared scene."""
    g = load_golden("render_v4_perturbed")
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m = m.to(DEV)
    d = sc.to(DEV)
    i = g["in"]
    ro, rd, near, far = (i[k].to(DEV) for k in ("rays_o", "rays_d", "near", "far"))
    t = O.draw_t_rand(ro.shape[0], 4)
    pr = torch.zeros(1, 3)
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    a = m.render(ro, rd, near, far, ps, None, None, None, None, None, None, d.intrs, d.c2ws, 1.0, None, t_rand=t, pts_random=pr)
    # other views: drop the last source view
    ids = [0, 1, 2, 3]
    imgs2, feats2, K2, c2 = d.imgs[ids], [f[ids] for f in d.features], d.intrs[ids], d.c2ws[ids]
    ps.set_views(imgs2, feats2, K2, c2)
    b = m.render(ro, rd, near, far, ps, None, None, None, None, None, None, K2, c2, 1.0, None, t_rand=t, pts_random=pr)
    fresh = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, imgs2, feats2, K2, c2)
    c = m.render(ro, rd, near, far, fresh, None, None, None, None, None, None, K2, c2, 1.0, None, t_rand=t, pts_random=pr)
    assert b["sampled_gray_val"].shape[0] == 3 and a["sampled_gray_val"].shape[0] == 4
    assert torch.equal(b["sampled_gray_val"], c["sampled_gray_val"]) and torch.equal(b["ref_gray_val"], c["ref_gray_val"])
    assert torch.equal(a["ref_gray_val"], b["ref_gray_val"])        # the reference view did not change


@pytest.mark.parametrize("mode", [_lib.MLP_TC, _lib.MLP_FFMA])
@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed"])
def test_second_order_smooth_vs_reference(name, mode):
    """SDFNetworkSparse.gradient's second return value (sdf_network.py:143-150, double autograd) by the analytic
    forward-over-reverse kernel, against the reference's own `smooth` at the reference's evaluated points, and
    smooth_error of render() against the golden."""
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m.mlp_mode = mode          # MLP_TC: the tcgen05 forward-over-reverse kernel; MLP_FFMA: the plain fp32 one
    m = m.to(DEV)
    d = sc.to(DEV)
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    pv = torch.from_numpy(g["out"]["_pts_valid"]).to(DEV)
    gr, sm = m.sdf_network.gradient(pv, ps)
    want = torch.as_tensor(g["out"]["_smooth_valid"])
    assert sm.shape == want.shape
    assert_close(sm, want, 2e-4, "smooth = Hessian . (1,1,1) vs reference double autograd")
    g32, sm2 = m.sdf_network.smooth(pv, ps, with_grad=True)
    assert torch.equal(sm, sm2)
    assert_close(g32, g["out"]["_grad_valid"], RTOL_FP32, "first-order gradient of the second-order kernel")
    # ragged sizes and the flag byte
    fl = torch.zeros(pv.shape[0], dtype=torch.uint8, device=DEV)
    fl[::2] = 2
    sm3 = m.sdf_network.smooth(pv, ps, flags=fl)
    assert torch.equal(sm3[::2], sm[::2]) and bool((sm3[1::2] == 0).all())
    assert m.sdf_network.smooth(pv[:13], ps).shape == (13, 3) and torch.equal(m.sdf_network.smooth(pv[:13], ps), sm[:13])
    # smooth_error of the full render
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = m.render(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), ps, None, None, None, None,
                   None, None, d.intrs, d.c2ws, 1.0, None)
    assert set(g["out"].keys()) - {k for k in g["out"] if k.startswith("_")} <= set(out.keys()), "all 18 reference keys"
    assert_close(out["smooth_error"], g["out"]["smooth_error"], 1e-3, "smooth_error")


def test_second_order_smooth_wild_points():
    g = load_golden("sdf_grid_24")
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m = m.to(DEV)
    d = sc.to(DEV)
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    wild = g["in"]["wild_pts"].to(DEV)
    sm = m.sdf_network.smooth(wild, ps)
    assert_close(sm, g["out"]["wild_smooth"], 2e-4, "smooth, out-of-range points and exact voxel centres")


def test_second_order_tensor_core_kernel_vs_fp32_kernel():
    """The tcgen05 edition against the plain fp32 kernel on 40 k points of the cfgD-shaped scene (many tiles per CTA,
    ragged last tile, masked-out points, points outside the volumes)."""
    from surf_b200 import synthetic
    import bench
    sc = synthetic.make_scene(5, 240, 320, 32, seed=12, device=DEV)
    m = bench.build_net(DEV)
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    g = torch.Generator(device=DEV).manual_seed(5)
    pts = (torch.rand(40000 + 77, 3, device=DEV, generator=g) * 2.4 - 1.2)
    fl = (torch.rand(pts.shape[0], device=DEV, generator=g) < 0.8).to(torch.uint8) * 2
    g_tc, s_tc = m.sdf_network.smooth(pts, ps, flags=fl, with_grad=True, mode=_lib.MLP_TC)
    g_32, s_32 = m.sdf_network.smooth(pts, ps, flags=fl, with_grad=True, mode=_lib.MLP_FFMA)
    assert bool(torch.isfinite(s_tc).all()) and float(s_32.abs().max()) > 0
    assert bool((s_tc[fl == 0] == 0).all()) and bool((g_tc[fl == 0] == 0).all())
    assert_close(g_tc, g_32, 1e-4, "first-order gradient, tensor-core vs fp32 second-order kernel")
    assert_close(s_tc, s_32, 2e-4, "smooth, tensor-core vs fp32 second-order kernel")
    # bitwise reproducible and independent of the tile a point lands in
    assert torch.equal(m.sdf_network.smooth(pts, ps, flags=fl, mode=_lib.MLP_TC), s_tc)
    assert torch.equal(m.sdf_network.smooth(pts[1000:3001], ps, flags=fl[1000:3001], mode=_lib.MLP_TC), s_tc[1000:3001])
    ps.destroy()
