"""tcgen05 building blocks (TMEM A operand, smem B operand, fp16 hi/lo split) against torch fp64."""
import ctypes as C

import pytest
import torch

from surf_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(K, N, split, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(128, K, generator=g) * 0.7).cuda()
    B = (torch.randn(N, K, generator=g) * 0.3).cuda()
    D = torch.zeros(128, N, device="cuda")
    _lib.check(_lib.load().surf_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, split,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tc_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    return float((D.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("K,N", [(16, 16), (32, 128), (160, 128), (128, 160), (64, 256)])
def test_tc_gemm_fp16(K, N):
    err = _run(K, N, 0)
    assert err < 2e-3, err          # single fp16 MMA: ~2^-11 relative


@pytest.mark.parametrize("K,N", [(32, 128), (160, 128), (128, 160)])
def test_tc_gemm_split_is_fp32_grade(K, N):
    err = _run(K, N, 1)
    assert err < 2e-6, err          # hi/lo split, 3 MMAs: ~2^-22 relative


# ------------------------------------------------------------------------------------------------
# tensor-core SDF MLP (forward) against the reference golden and the fp32 FFMA kernel
# ------------------------------------------------------------------------------------------------
from helpers import RTOL_FP32, assert_close, load_golden, scene_from_recipe  # noqa: E402
from surf_b200 import conf  # noqa: E402
from surf_b200.modules.implicit_surface import ImplicitSurface  # noqa: E402


# mode 1: the shipped tensor-core kernels (pipelined one-tile kernel, sdf_tc2.cu); mode 3: the first-generation ones
@pytest.fixture(params=[1, 3, 5])
def tc_mode(request):
    _lib.set_mlp_mode(request.param)
    yield request.param
    _lib.set_mlp_mode(0)


@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed"])
def test_tc_sdf_forward_vs_reference(name, tc_mode):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"]).to("cuda")
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m = m.cuda()
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    pv = torch.from_numpy(g["out"]["_pts_valid"]).cuda()
    sdf_tc = m.sdf_network.sdf(pv, ps)
    assert_close(sdf_tc, g["out"]["_sdf_full"][:, :1], RTOL_FP32, "tensor-core sdf vs reference golden")
    _lib.set_mlp_mode(0)
    sdf_ffma = m.sdf_network.sdf(pv, ps)
    _lib.set_mlp_mode(tc_mode)
    assert float((sdf_tc - sdf_ffma).abs().max()) < 2e-5
    for n in (1, 127, 128, 129, 255, 257, 1000):
        assert torch.equal(m.sdf_network.sdf(pv[:n], ps), sdf_tc[:n])


def test_tc_sdf_grid_vs_reference(tc_mode):
    g = load_golden("sdf_grid_24")
    sc = scene_from_recipe(g["recipe"]).to("cuda")
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m = m.cuda()
    ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
    u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 24)
    assert_close(u, g["out"]["u"], RTOL_FP32, "tensor-core u grid vs reference golden")
    wild = g["in"]["wild_pts"].cuda()
    assert_close(m.sdf_network.sdf(wild, ps), g["out"]["wild_full"][:, :1], RTOL_FP32, "tc sdf, out-of-range points")


# ------------------------------------------------------------------------------------------------
# one-tile tensor-core kernel: forward + analytic gradient
# ------------------------------------------------------------------------------------------------
def _setup(name):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m = m.cuda()
    d = sc.to("cuda")
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    return g, sc, d, m, ps


@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed"])
def test_tc_gradient_vs_reference(name, tc_mode):
    g, sc, d, m, ps = _setup(name)
    pv = torch.from_numpy(g["out"]["_pts_valid"]).cuda()
    sdf, grad = m.sdf_network.gradient(pv, ps, with_sdf=True)
    assert_close(sdf, g["out"]["_sdf_full"][:, :1], RTOL_FP32, "tc sdf (gradient kernel) vs reference golden")
    assert_close(grad, g["out"]["_grad_valid"], RTOL_FP32, "tc d sdf / d x vs reference autograd")
    s2, g2 = m.sdf_network.gradient(pv, ps, with_sdf=True)
    assert torch.equal(sdf, s2) and torch.equal(grad, g2), "tensor-core gradient kernel must be deterministic"
    for n in (1, 127, 129, 300):
        s3, g3 = m.sdf_network.gradient(pv[:n], ps, with_sdf=True)
        assert torch.equal(s3, sdf[:n]) and torch.equal(g3, grad[:n])


def test_tc_gradient_wild_points(tc_mode):
    g, sc, d, m, ps = _setup("sdf_grid_24")
    wild = g["in"]["wild_pts"].cuda()
    s, gr = m.sdf_network.gradient(wild, ps, with_sdf=True)
    assert_close(s, g["out"]["wild_full"][:, :1], RTOL_FP32, "tc sdf, out-of-range points")
    assert_close(gr, g["out"]["wild_grad"], RTOL_FP32, "tc gradient, out-of-range points")


@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed", "render_miss"])
def test_tc_render_end_to_end(name, tc_mode):
    import surf_oracle as O
    from helpers import assert_equal_int
    g, sc, d, m, ps = _setup(name)
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    t_rand = O.draw_t_rand(i["rays_o"].shape[0], 4)
    pts_random = torch.rand(1024, 3) * 2 - 1
    out = m.render(i["rays_o"].cuda(), i["rays_d"].cuda(), i["near"].cuda(), i["far"].cuda(), ps, None, None, None, None,
                   None, None, None, None, 1.0, None, t_rand=t_rand, pts_random=pts_random, return_stages=True)
    net = O.OracleNet(g["sd"])
    ref = O.render(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, sc.volumes, sc.sparse_idxes,
                   sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0, t_rand=t_rand, pts_random=pts_random,
                   return_stages=True)
    mism = int(((out["_point_flags"].cpu() & 1).bool() != ref["_voxel_mask"]).sum())
    assert mism <= 2
    assert_close(out["gradients"], ref["gradients"], RTOL_FP32, "gradients")
    assert_close(out["sparse_sdf"], ref["sparse_sdf"], RTOL_FP32, "sparse_sdf")
    if mism == 0 and name != "render_v4_perturbed":
        for k in ("color_fine", "render_depth", "sdf_depth", "normal", "weight_sum"):
            assert_close(out[k], ref[k], 5e-4, k, floor=1e-2)
        assert_equal_int(out["valid_mask"], ref["valid_mask"], "valid_mask")


@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed"])
def test_tc_blend_vs_reference(name, tc_mode):
    import surf_oracle as O
    from helpers import blend_envelope
    g = load_golden(name)
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m = m.cuda()
    o = g["out"]
    net = O.OracleNet(g["sd"])
    fv, rd, mk = (torch.from_numpy(o[k]) for k in ("_feat_views", "_ray_diff", "_view_mask"))
    got = m.color_network(fv.cuda(), rd.cuda(), mk.cuda()).cpu()
    ref = torch.from_numpy(o["_blend_rgb"])
    _, env = blend_envelope(O, net, fv, rd, mk)
    err = (got - ref).abs().max(dim=1)[0]
    tol = RTOL_FP32 * float(ref.abs().max()) + 2.0 * env
    assert bool((err <= tol).all()), "tc blend rgb: %d/%d points beyond 1e-4 + envelope (max err %.3e)" % (
        int((err > tol).sum()), err.numel(), float(err.max()))
    mk0 = torch.zeros_like(mk)
    assert_close(m.color_network(fv.cuda(), rd.cuda(), mk0.cuda()), O.blend(net, fv, rd, mk0), RTOL_FP32,
                 "tc blend rgb, nothing visible")
    again = m.color_network(fv.cuda(), rd.cuda(), mk.cuda()).cpu()
    assert torch.equal(got, again)


# ------------------------------------------------------------------------------------------------
# opt-in reduced-precision mode (north_star: "1e-2 relative for the opt-in bf16 MLP mode"): mode 4 issues one
# fp16 MMA per product instead of three (fp16 rather than bf16: same cost, ~8x smaller error)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init"])
def test_fast_mode_within_1e2(name):
    import surf_oracle as O
    from helpers import RTOL_BF16
    _lib.set_mlp_mode(4)
    try:
        g, sc, d, m, ps = _setup(name)
        pv = torch.from_numpy(g["out"]["_pts_valid"]).cuda()
        sdf, grad = m.sdf_network.gradient(pv, ps, with_sdf=True)
        assert_close(sdf, g["out"]["_sdf_full"][:, :1], RTOL_BF16, "fast-mode sdf")
        assert_close(grad, g["out"]["_grad_valid"], RTOL_BF16, "fast-mode gradient")
        assert_close(m.sdf_network.sdf(pv, ps), g["out"]["_sdf_full"][:, :1], RTOL_BF16, "fast-mode sdf (forward kernel)")
        o = g["out"]
        fv, rd, mk = (torch.from_numpy(o[k]) for k in ("_feat_views", "_ray_diff", "_view_mask"))
        rgb = m.color_network(fv.cuda(), rd.cuda(), mk.cuda())
        assert_close(rgb, o["_blend_rgb"], RTOL_BF16, "fast-mode blend rgb")
        err = float((sdf.cpu() - torch.from_numpy(o["_sdf_full"][:, :1])).abs().max())
        assert err > 1e-6, "fast mode should differ measurably from the fp32-grade path"
    finally:
        _lib.set_mlp_mode(0)
