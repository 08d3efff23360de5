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


@pytest.fixture
def tc_mode():
    _lib.set_mlp_mode(1)
    yield
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
    _lib.set_mlp_mode(1)
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
