"""MatchingField on the GPU (csrc/matching.cu) against the unmodified reference's outputs (tests/golden/
matching_field.npz) and the oracle: every view, stage 0 / 1 / 3, with and without jitter."""
import pytest
import torch

import surf_oracle as O
from helpers import RTOL_FP32, assert_close, load_golden, scene_from_recipe
from surf_b200 import conf
from surf_b200.modules.matching_field import MatchingField

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mf():
    c = conf.ConfigTree()
    c.put("n_samples_depths", [128, 64, 32, 16])
    c.put("n_importance_depths", [128, 64, 32, 16])
    c.put("up_sample_steps", [4, 4, 4, 4])
    c.put("depth_res_levels", [4, 2, 2, 1])
    return MatchingField(c)


def test_depth_maps_vs_reference():
    g = load_golden("matching_field")
    sc = scene_from_recipe(g["recipe"])
    d = sc.to(DEV)
    mf = _mf()
    rr = [1.0, 0.4, 0.1, 0.01]
    ipts = {"near_fars": g["in"]["near_fars"].to(DEV), "c2ws": d.c2ws, "intrs": d.intrs, "imgs": d.imgs, "src_idx": 1}
    d0, o0 = mf(ipts, d.matching_volume, 0, rr, None)
    assert len(d0) == 3 and d0[0].shape == (48, 64) and d0[0].device.type == "cuda"
    # later stages are fed the REFERENCE's previous maps, so that each stage is compared in isolation
    ref0 = [t.to(DEV) for t in torch.as_tensor(g["out"]["depth_s0"])]
    ref1 = [t.to(DEV) for t in torch.as_tensor(g["out"]["depth_s1"])]
    d1, o1 = mf(ipts, d.matching_volume, 1, rr, ref0)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    d1p, o1p = mf(ipts, d.matching_volume, 1, rr, ref0, perturb=True)
    d3, o3 = mf(ipts, d.matching_volume, 3, rr, ref1)
    for tag, (dd, oo) in {"s0": (d0, o0), "s1": (d1, o1), "s1p": (d1p, o1p), "s3": (d3, o3)}.items():
        assert_close(torch.stack(dd), g["out"]["depth_" + tag], RTOL_FP32, "depth " + tag)
        assert_close(torch.stack(oo), g["out"]["occ_" + tag], RTOL_FP32, "occ_reg " + tag)
    # chained like build_volumes does (its own stage-0 maps in): still within 1e-4 of the reference
    d1c, _ = mf(ipts, d.matching_volume, 1, rr, d0)
    assert_close(torch.stack(d1c), g["out"]["depth_s1"], 2 * RTOL_FP32, "depth s1, chained")
    # determinism
    d0b, _ = mf(ipts, d.matching_volume, 0, rr, None)
    assert all(torch.equal(a, b) for a, b in zip(d0, d0b))
