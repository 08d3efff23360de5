"""Volume.* on the GPU (csrc/volume.cu) against the unmodified reference's outputs (tests/golden/volume.npz), and the
compact scene builder (surf_scene_create_sparse) against the scene prepared from the reference-layout tensors."""
import pytest
import torch

import surf_oracle as O
from helpers import RTOL_FP32, assert_close, assert_equal_int, load_golden, scene_from_recipe
from surf_b200 import conf, synthetic
from surf_b200.modules import projector as P
from surf_b200.modules.matching_field import MatchingField
from surf_b200.modules.volume import Volume

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _volume(g, base):
    c = conf.ConfigTree()
    c.put("base_volume_dim", [base, base, base])
    v = Volume(c)
    v.load_state_dict({k: torch.as_tensor(t) for k, t in g["sd"].items()}, strict=True)
    return v.to(DEV)


def test_volume_producers_vs_reference():
    g = load_golden("volume")
    sc = scene_from_recipe(g["recipe"])
    d = sc.to(DEV)
    o = {k: torch.as_tensor(v) for k, v in g["out"].items()}
    base = int(g["recipe"]["base"])
    vol = _volume(g, base)
    feats = d.features[::-1]                                   # coarse -> fine
    c0 = vol.init_coords().to(DEV)
    assert torch.equal(c0.cpu(), o["c0"])
    fv0, m0 = vol.back_proj_multiscale(feats, c0, d.intrs, d.c2ws, 0)
    assert_equal_int(m0, o["m0"], "frustum mask, stage 0")
    assert_close(fv0, o["fv0"], RTOL_FP32, "back-projected features, stage 0")
    c0m = c0[m0]
    reg0 = o["reg0"].to(DEV)
    mv0, mk0 = vol.sparse2dense(reg0[:, :1], c0m, None)
    assert torch.equal(mv0.cpu(), o["mv0"]) and torch.equal(mk0.cpu(), o["mk0"])
    assert torch.equal(vol.get_index(c0m).cpu(), o["idx0"])
    c1, f1 = vol.up_sample(c0m.clone(), reg0)
    assert torch.equal(c1.cpu(), o["c1"]) and torch.equal(f1.cpu(), o["f1"])
    c1f, f1f = vol.depth_filtering(list(o["depths"].to(DEV)), c1, f1, d.intrs, d.c2ws, 0.4)
    assert torch.equal(c1f.cpu(), o["c1f"]) and torch.equal(f1f.cpu(), o["f1f"]), "depth-consistency filter"
    fv1, m1 = vol.back_proj_multiscale(feats, c1f, d.intrs, d.c2ws, 1)
    assert_equal_int(m1, o["m1"], "frustum mask, stage 1")
    assert_close(fv1, o["fv1"], RTOL_FP32, "back-projected features, stage 1")
    c1m = c1f[m1]
    reg1 = o["reg1"].to(DEV)
    mv1, mk1 = vol.sparse2dense(reg1[:, :1], c1m, mv0)
    assert_close(mv1, o["mv1"], 1e-6, "matching volume over the up-sampled coarser one")
    assert torch.equal(mk1.cpu(), o["mk1"]) and torch.equal(vol.get_index(c1m).cpu(), o["idx1"])
    # empty input
    e_fv, e_m = vol.back_proj_multiscale(feats, c1f[:0], d.intrs, d.c2ws, 1)
    assert e_fv.shape == (0, 8) and e_m.shape == (0,)


def test_compact_scene_equals_the_reference_layout_scene():
    """surf_scene_create_sparse (coordinates in, compact layout out) == surf_scene_create on the reference-layout
    tensors the Volume methods return: same sparse gather, same voxel masks, same matching-volume probes, bit for bit."""
    g = load_golden("volume")
    o = {k: torch.as_tensor(v).to(DEV) for k, v in g["out"].items()}
    base = int(g["recipe"]["base"])
    c0m, c1m = o["c0"][o["m0"]], o["c1f"][o["m1"]]
    reg0, reg1 = o["reg0"], o["reg1"]
    comp = Volume.to_prepared_scene([c0m, c1m], [reg0[:, 1:], reg1[:, 1:]], [reg0[:, :1], reg1[:, :1]], [base, 2 * base])
    # reference layouts, RENDERER order (fine -> coarse)
    ref = P.PreparedScene([reg1[:, 1:].contiguous(), reg0[:, 1:].contiguous()], [o["idx1"], o["idx0"]],
                          [o["mk1"], o["mk0"]], o["mv1"])
    gen = torch.Generator().manual_seed(0)
    pts = (torch.rand(20000, 3, generator=gen) * 2.4 - 1.2).to(DEV)
    assert torch.equal(P.lookup_sparse_volume(pts, comp), P.lookup_sparse_volume(pts, ref))
    assert torch.equal(P.lookup_volume(pts, comp, "nearest"), P.lookup_volume(pts, ref, "nearest"))
    # the matching volume through the depth-map kernel of both scenes
    sc = scene_from_recipe(g["recipe"]).to(DEV)
    cf = conf.ConfigTree()
    for k, v in (("n_samples_depths", [128, 64, 32, 16]), ("n_importance_depths", [128, 64, 32, 16]),
                 ("up_sample_steps", [4, 4, 4, 4]), ("depth_res_levels", [4, 2, 2, 1])):
        cf.put(k, v)
    mf = MatchingField(cf)
    ipts = {"near_fars": torch.tensor([[float(sc.near), float(sc.far)]] * 3), "c2ws": sc.c2ws, "intrs": sc.intrs, "imgs": sc.imgs}
    da, _ = mf(ipts, comp, 0, [1.0, 0.4, 0.1, 0.01])
    db, _ = mf(ipts, ref, 0, [1.0, 0.4, 0.1, 0.01])
    # (the compact builder up-samples the coarser matching volume with its own kernel: 1e-6 of ATen's F.interpolate)
    for a, b in zip(da, db):
        assert_close(a, b, 1e-5, "depth map through the compact scene's matching volume")
    st = comp.stats()
    assert st["n_vox"] == [int(c1m.shape[0]), int(c0m.shape[0])]
