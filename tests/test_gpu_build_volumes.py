"""SuRF.build_volumes / init_volumes (surf.py:65-131) on the kernels of FeatureNetwork / Volume / MatchingField, against
the UNMODIFIED reference's own build_volumes run with the same stand-in regulariser (tests/golden/build_volumes.npz;
the torchsparse network is the one piece that is not built, oracle/standin_reg.py takes its place on both sides)."""
import numpy as np
import pytest
import torch

import standin_reg
from helpers import assert_close, load_golden, scene_from_recipe
from surf_b200 import conf
from surf_b200.surf import SuRF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(base):
    c = conf.parse_string("""
        range_ratios = [1.0, 0.4, 0.1, 0.01]
        feature_network { d_in = 3, d_base = 8, d_out = [4, 4, 4, 4] }
        volume { base_volume_dim = [%d, %d, %d] }
        matching_field { n_samples_depths = [128, 64, 32, 16], n_importance_depths = [128, 64, 32, 16],
                         up_sample_steps = [4, 4, 4, 4], depth_res_levels = [4, 2, 2, 1] }
    """ % (base, base, base))
    c.put("implicit_surface", conf.default_implicit_surface_conf())
    return SuRF(c)


def _dense(volume, idx):
    """(n, c) rows + index table -> (D, H, W, c) with NaN where empty."""
    out = torch.full(tuple(idx.shape) + (volume.shape[1],), float("nan"))
    occ = idx >= 0
    out[occ] = volume[idx[occ].long()]
    return out


def test_build_volumes_vs_the_reference():
    g = load_golden("build_volumes")
    sc = scene_from_recipe(g["recipe"])
    base = int(g["recipe"]["base"])
    model = _model(base)
    model.feature_network.load_state_dict({k[len("feature_network."):]: v for k, v in g["sd"].items()
                                           if k.startswith("feature_network.")}, strict=True)
    model.volume.load_state_dict({k[len("volume."):]: v for k, v in g["sd"].items() if k.startswith("volume.")}, strict=True)
    model = model.to(DEV)
    with pytest.raises(NotImplementedError):
        model.build_volumes({}, [])                      # no regularisation network plugged in
    model.reg_network = standin_reg.StandinReg()
    d = sc.to(DEV)
    near_fars = torch.stack([torch.tensor([float(sc.near), float(sc.far)])] * sc.nv)
    ipts = {"imgs": d.imgs, "intrs": d.intrs, "c2ws": d.c2ws, "near": d.near, "far": d.far, "near_fars": near_fars, "src_idx": 1}
    features = [g["in"]["features%d" % i].to(DEV) for i in range(4)]      # the reference's own pyramid (coarse -> fine)
    outputs, volumes, idxs, masks, matching = model.build_volumes(ipts, features, False)
    o = g["out"]
    assert len(volumes) == len(idxs) == len(masks) == 4
    for s in range(4):
        want_idx = torch.from_numpy(o["sparse_idx%d" % s]).long()
        got_idx = idxs[s].cpu()
        assert got_idx.dtype == torch.int64 and got_idx.shape == want_idx.shape == (base * 2 ** s,) * 3
        occ_g, occ_w = got_idx >= 0, want_idx >= 0
        differ = int((occ_g != occ_w).sum())
        # the occupied sets follow from threshold decisions on depths that carry the kernels' 1e-6 differences
        assert differ <= max(2, int(0.002 * int(occ_w.sum()))), "stage %d: %d voxels differ in occupancy" % (s, differ)
        assert torch.equal(masks[s].cpu()[0, 0] > 0, occ_g)
        both = occ_g & occ_w
        dv_g, dv_w = _dense(volumes[s].cpu(), got_idx), _dense(torch.from_numpy(o["volume%d" % s]), want_idx)
        assert volumes[s].shape[1] == 7
        assert_close(dv_g[both], dv_w[both], 2e-4, "stage %d feature volume on the common voxels" % s)
        assert_close(outputs["depth_stage%d" % s], o["depth_stage%d" % s], 2e-4, "depth_stage%d" % s)
        assert_close(outputs["depth_src_stage%d" % s], o["depth_src_stage%d" % s], 2e-4, "depth_src_stage%d" % s)
        print("stage %d: %d voxels, %d occupancy differences" % (s, int(occ_w.sum()), differ))
    mv_g, mv_w = matching.cpu(), torch.from_numpy(o["matching_volume"])
    close = (mv_g - mv_w).abs() <= 2e-4 * float(mv_w.abs().max())
    assert float(close.float().mean()) > 0.998, "matching volume"


def test_init_volumes_then_render():
    """init_volumes (surf.py:65-78) end to end on the GPU pieces: images -> pyramid -> volumes -> a rendered batch."""
    g = load_golden("build_volumes")
    sc = scene_from_recipe(g["recipe"])
    model = _model(int(g["recipe"]["base"])).to(DEV)
    model.reg_network = standin_reg.StandinReg()
    d = sc.to(DEV)
    near_fars = torch.stack([torch.tensor([float(sc.near), float(sc.far)])] * sc.nv)
    ipts = {"imgs": d.imgs, "intrs": d.intrs, "c2ws": d.c2ws, "near": d.near, "far": d.far, "near_fars": near_fars, "src_idx": 1}
    model.init_volumes(ipts)
    assert model.has_vol and len(model.volumes) == 4 and len(model.features) == 4
    from surf_b200 import synthetic
    o, dd = synthetic.random_pixel_rays(sc, 64, seed=3)
    ipts.update({"rays_o": o.to(DEV), "rays_d": dd.to(DEV)})
    torch.manual_seed(0)
    out = model("train", ipts)
    assert out["color_fine"].shape == (64, 3) and bool(torch.isfinite(out["color_fine"]).all())


def test_forward_without_volumes_rebuilds_them_per_call():
    """The generalisable branch of SuRF.forward (surf.py:137-148): pyramid + volumes from ipts["imgs"] on every call,
    the matching-field depth maps in the outputs; same render as init_volumes() followed by the has_vol branch."""
    g = load_golden("build_volumes")
    sc = scene_from_recipe(g["recipe"])
    model = _model(int(g["recipe"]["base"])).to(DEV)
    model.reg_network = standin_reg.StandinReg()
    model.match_feature_network.load_state_dict(model.feature_network.state_dict())
    d = sc.to(DEV)
    near_fars = torch.stack([torch.tensor([float(sc.near), float(sc.far)])] * sc.nv)
    from surf_b200 import synthetic
    o, dd = synthetic.random_pixel_rays(sc, 96, seed=4)
    ipts = {"imgs": d.imgs, "intrs": d.intrs, "c2ws": d.c2ws, "near": d.near, "far": d.far, "near_fars": near_fars,
            "src_idx": 1, "rays_o": o.to(DEV), "rays_d": dd.to(DEV)}
    torch.manual_seed(1)
    a = model("test", ipts)         # any mode but "train": the matching field is not jittered (surf.py:139)
    assert not model.has_vol
    assert all(("depth_stage%d" % s) in a and ("depth_src_stage%d" % s) in a for s in range(4))
    assert tuple(a["depth_stage3"].shape) == (sc.H, sc.W)
    model.init_volumes(ipts)
    torch.manual_seed(1)
    b = model("test", ipts)
    for k in ("color_fine", "render_depth", "sdf_depth"):
        assert torch.equal(torch.as_tensor(a[k]), torch.as_tensor(b[k])), k


def test_compact_init_volumes_renders_the_same():
    """init_volumes(compact=True): the scene goes straight into the prepared layout (no int64 tables / fp32 masks) and
    renders bit-identically to the reference-layout path."""
    g = load_golden("build_volumes")
    sc = scene_from_recipe(g["recipe"])
    d = sc.to(DEV)
    near_fars = torch.stack([torch.tensor([float(sc.near), float(sc.far)])] * sc.nv)
    from surf_b200 import synthetic
    o, dd = synthetic.random_pixel_rays(sc, 96, seed=5)
    ipts = {"imgs": d.imgs, "intrs": d.intrs, "c2ws": d.c2ws, "near": d.near, "far": d.far, "near_fars": near_fars,
            "src_idx": 1, "rays_o": o.to(DEV), "rays_d": dd.to(DEV)}
    outs = []
    for compact in (False, True):
        torch.manual_seed(11)
        model = _model(int(g["recipe"]["base"])).to(DEV)
        model.reg_network = standin_reg.StandinReg()
        model.init_volumes(ipts, compact=compact)
        assert model.has_vol and (model.prepared is not None) == compact
        if compact:
            assert model.volumes is None and model.sparse_idxes is None
        torch.manual_seed(2)
        outs.append(model("train", ipts))
    for k in ("color_fine", "render_depth", "sdf_depth", "gradients", "weights"):
        assert torch.equal(outs[0][k], outs[1][k]), k
