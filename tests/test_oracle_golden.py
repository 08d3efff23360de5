"""Pins oracle/surf_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/make_golden.py).  Integers bit-exact; floats <= 1e-5 scale-relative (the oracle
uses the same ATen ops, differences are summation-order only)."""
import numpy as np
import pytest
import torch

import surf_oracle as O
from helpers import assert_close, assert_equal_int, load_golden, scene_from_recipe

TOL = 1e-5
RENDER_CASES = ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed", "render_miss"]


def _net(g):
    return O.OracleNet(g["sd"])


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_matches_reference(name):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    net = _net(g)
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = O.render(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, sc.volumes,
                   sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0,
                   return_stages=True)
    ref = g["out"]
    assert_equal_int(out["_voxel_mask"], ref["_voxel_mask"], "voxel_mask")
    assert_equal_int(out["valid_mask"], ref["valid_mask"], "valid_mask")
    assert_equal_int(out["inside_sphere"], ref["inside_sphere"], "inside_sphere")
    assert_equal_int(out["mid_inside_sphere"], ref["mid_inside_sphere"], "mid_inside_sphere")
    assert_equal_int(out["_view_mask"][out["_compute_mask"]], ref["_view_mask"], "view_mask")
    assert np.array_equal(out["mid_z_vals"].numpy(), ref["mid_z_vals"]), "mid_z_vals must be bit-identical"
    for k in ["color_fine", "render_depth", "sdf_depth", "normal", "gradients", "weights", "weight_sum",
              "weight_max", "sparse_sdf", "s_val", "gradient_error"]:
        assert_close(out[k], ref[k], TOL, k)


@pytest.mark.parametrize("name", RENDER_CASES)
def test_stage_functions_match_reference(name):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    net = _net(g)
    ref = g["out"]
    pv = torch.from_numpy(ref["_pts_valid"])
    feats = O.lookup_sparse(pv, sc.volumes, sc.sparse_idxes)
    assert_close(feats, ref["_sparse_feats"], 1e-6, "sparse feats")
    fv, rd, mv = O.lookup_feature(pv, sc.imgs, sc.intrs, sc.c2ws, sc.features)
    assert_equal_int(mv, ref["_view_mask"], "view mask")
    assert_close(fv, ref["_feat_views"], 1e-6, "feat_views")
    assert_close(rd, ref["_ray_diff"], 1e-6, "ray_diff")
    rgb = O.blend(net, torch.from_numpy(ref["_feat_views"]), torch.from_numpy(ref["_ray_diff"]),
                  torch.from_numpy(ref["_view_mask"]))
    assert_close(rgb, ref["_blend_rgb"], TOL, "blend rgb")
    full = O.sdf_forward(net, pv, sc.volumes, sc.sparse_idxes)
    assert_close(full, ref["_sdf_full"], TOL, "sdf full output")
    s, gr = O.sdf_gradient(net, pv, sc.volumes, sc.sparse_idxes)
    assert_close(gr, ref["_grad_valid"], TOL, "sdf gradient (autograd)")
    s2, gr2 = O.sdf_gradient_analytic(net, pv, sc.volumes, sc.sparse_idxes)
    assert_close(gr2, ref["_grad_valid"], 2e-5, "sdf gradient (analytic reverse pass)")
    assert_close(s2, ref["_sdf_full"][:, :1], TOL, "sdf (analytic path)")


def test_validate_image_matches_reference():
    g = load_golden("validate_24x32")
    sc = scene_from_recipe(g["recipe"])
    net = _net(g)
    i = g["in"]
    r = int(g["recipe"]["res_level"])
    hw = (sc.H // r, sc.W // r)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = O.validate_image(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, sc.volumes,
                           sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, hw)
    for k in ["color_fine", "img_fine", "normal_img", "sdf_depth", "render_depth"]:
        assert_close(out[k], g["out"][k], TOL, k)


def test_sdf_grid_matches_reference():
    g = load_golden("sdf_grid_24")
    sc = scene_from_recipe(g["recipe"])
    net = _net(g)
    res = int(g["recipe"]["resolution"])
    u = O.sdf_grid(net, sc.volumes, sc.sparse_idxes, [-1, -1, -1], [1, 1, 1], res)
    assert_close(u, g["out"]["u"], TOL, "u grid")
    sub = O.sdf_grid(net, sc.volumes, sc.sparse_idxes, [-1, -1, -1], [1, 1, 1], res, x_range=(5, 11),
                     z_range=(3, 20))
    assert_close(sub, g["out"]["u"][5:11, :, 3:20], TOL, "u sub-box")
    wild = g["in"]["wild_pts"]
    full = O.sdf_forward(net, wild, sc.volumes, sc.sparse_idxes)
    assert_close(full, g["out"]["wild_full"], TOL, "out-of-range points")
    _, gr = O.sdf_gradient(net, wild, sc.volumes, sc.sparse_idxes)
    assert_close(gr, g["out"]["wild_grad"], TOL, "out-of-range gradient")
    _, gr2 = O.sdf_gradient_analytic(net, wild, sc.volumes, sc.sparse_idxes)
    assert_close(gr2, g["out"]["wild_grad"], 2e-5, "out-of-range gradient (analytic)")


def test_reference_gradient_moves_with_the_last_bit_of_the_position():
    """The conditioning fact behind helpers.explain_gradient_mismatches: a ONE-ulp shift of the sample position changes
    the reference's own d sdf / d x by an amount that grows with the volume resolution (trilinear features: derivative
    changes at (feature difference) / voxel^2).  Measured here on the oracle at two resolutions."""
    import bench              # repo root is on sys.path (tests/conftest.py)
    from surf_b200 import synthetic
    m = bench.build_net(None, seed=0)
    net = O.OracleNet({k: v.detach() for k, v in m.state_dict().items()})
    dev = {}
    for base in (4, 32):
        sc = synthetic.make_scene(3, 48, 64, base, seed=3, device="cpu")
        g = torch.Generator().manual_seed(1)
        pts = torch.nn.functional.normalize(torch.randn(4000, 3, generator=g), dim=1) * (0.5 + 0.02 * torch.randn(4000, 1, generator=g))
        _, g0 = O.sdf_gradient(net, pts, sc.volumes, sc.sparse_idxes)
        _, g1 = O.sdf_gradient(net, torch.nextafter(pts, torch.full_like(pts, 10.0)), sc.volumes, sc.sparse_idxes)
        on_face = O.voxel_face_distance(pts, sc.sparse_idxes) < 2.5
        d = ((g1 - g0).abs().max(dim=1)[0] / g0.abs().max())[~on_face]
        dev[base] = float(torch.quantile(d, 0.999))
    assert dev[32] > 3 * dev[4], dev                  # 8x the resolution: the sensitivity grows with it
    assert dev[32] > 3e-6, dev                        # already >= 3 % of the 1e-4 budget per ulp at 256^3 (704^3: x2.75)


@pytest.mark.parametrize("name", ["render_v2_perturbed", "render_v2_init", "render_v4_perturbed", "render_miss"])
def test_training_extras_match_the_reference(name):
    """The training-only outputs of render() (implicit_surface.py:172, 218-245; projector.py:560-645): smooth_error
    (second-order autograd) and the warped 11x11 feature patches, restated in the oracle, against the reference's own
    outputs."""
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    net = O.OracleNet(g["sd"])
    i = g["in"]
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = O.render(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, sc.volumes, sc.sparse_idxes,
                   sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0, extras=True)
    for k in ("ref_gray_val", "sampled_gray_val"):
        assert out[k].shape == tuple(g["out"][k].shape)
        assert_close(out[k], g["out"][k], 1e-6, k)
    assert_close(out["smooth_error"], g["out"]["smooth_error"], 1e-5, "smooth_error")


def test_matching_field_matches_the_reference():
    """MatchingField.forward (matching_field.py:74-141) restated in the oracle vs the unmodified reference: stage 0,
    stage 1 (two windows around the stage-0 depth), a jittered pass, stage 3 at full resolution."""
    g = load_golden("matching_field")
    sc = scene_from_recipe(g["recipe"])
    ipts = {"near_fars": g["in"]["near_fars"], "c2ws": sc.c2ws, "intrs": sc.intrs, "imgs": sc.imgs, "src_idx": 1}
    rr, ns, lv = [1.0, 0.4, 0.1, 0.01], [128, 64, 32, 16], [4, 2, 2, 1]
    d0, o0 = O.matching_field_forward(ns, lv, ipts, sc.matching_volume, 0, rr, None)
    d1, o1 = O.matching_field_forward(ns, lv, ipts, sc.matching_volume, 1, rr, d0)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    d1p, o1p = O.matching_field_forward(ns, lv, ipts, sc.matching_volume, 1, rr, d0, perturb=True)
    d3, o3 = O.matching_field_forward(ns, lv, ipts, sc.matching_volume, 3, rr, d1)
    for tag, (dd, oo) in {"s0": (d0, o0), "s1": (d1, o1), "s1p": (d1p, o1p), "s3": (d3, o3)}.items():
        assert_close(torch.stack(dd), g["out"]["depth_" + tag], 1e-6, "depth " + tag)
        assert_close(torch.stack(oo), g["out"]["occ_" + tag], 1e-6, "occ_reg " + tag)


def test_volume_producers_match_the_reference():
    """Volume.* (volume.py:21-168) restated in the oracle vs the unmodified reference."""
    g = load_golden("volume")
    sc = scene_from_recipe(g["recipe"])
    o = {k: torch.as_tensor(v) for k, v in g["out"].items()}
    agg = {k[len("agg_mlp."):]: torch.as_tensor(v) for k, v in g["sd"].items()}
    feats = sc.features[::-1]
    base = int(g["recipe"]["base"])
    vs0, org = O.volume_voxel_size([base] * 3)
    vs1, _ = O.volume_voxel_size([2 * base] * 3)
    c0 = O.volume_init_coords([base] * 3)
    assert torch.equal(c0, o["c0"])
    fv0, m0 = O.volume_back_proj(agg, feats, c0, vs0, org, sc.intrs, sc.c2ws, 0)
    assert torch.equal(m0, o["m0"]) and torch.equal(fv0, o["fv0"])
    c0m = c0[m0]
    mv0, mk0 = O.volume_sparse2dense(o["reg0"][:, :1], c0m, [base] * 3)
    assert torch.equal(mv0, o["mv0"]) and torch.equal(mk0, o["mk0"])
    assert torch.equal(O.volume_get_index(c0m, [base] * 3), o["idx0"])
    c1, f1 = O.volume_up_sample(c0m, o["reg0"])
    assert torch.equal(c1, o["c1"]) and torch.equal(f1, o["f1"])
    dm = O.volume_depth_filter_mask(list(o["depths"]), c1, vs1, org, sc.intrs, sc.c2ws, 0.4)
    assert torch.equal(c1[dm], o["c1f"])
    fv1, m1 = O.volume_back_proj(agg, feats, c1[dm], vs1, org, sc.intrs, sc.c2ws, 1)
    assert torch.equal(m1, o["m1"]) and torch.equal(fv1, o["fv1"])
    mv1, mk1 = O.volume_sparse2dense(o["reg1"][:, :1], c1[dm][m1], [2 * base] * 3, mv0)
    assert torch.equal(mv1, o["mv1"]) and torch.equal(mk1, o["mk1"])
    assert torch.equal(O.volume_get_index(c1[dm][m1], [2 * base] * 3), o["idx1"])


def test_fpn_oracle_equals_the_reference_feature_network():
    """oracle/fpn_oracle.py against the unmodified reference FeatureNetwork (tests/golden/fpn.npz)."""
    import fpn_oracle
    g = load_golden("fpn")
    outs = fpn_oracle.feature_network_forward(g["sd"], g["in"]["imgs"])
    assert len(outs) == 4
    for i, o in enumerate(outs):
        want = torch.from_numpy(g["out"]["feat%d" % i])
        assert o.shape == want.shape == (2, 4, 40 // 2 ** (3 - i), 56 // 2 ** (3 - i))
        assert torch.equal(o, want), "stage %d differs from the reference" % i
