"""The 2-D feature pyramid kernels (csrc/fpn.cu; FeatureNetwork, models/modules/feature_network.py:126-178) against the
reference's own outputs (tests/golden/fpn.npz) and the torch restatement at other sizes."""
import numpy as np
import pytest
import torch

import fpn_oracle
from helpers import assert_close, load_golden
from surf_b200 import conf
from surf_b200.modules.feature_network import FeatureNetwork

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _conf():
    return conf.parse_string("d_in = 3\nd_base = 8\nd_out = [4, 4, 4, 4]")


def test_reference_state_dict_loads_and_outputs_match_the_golden():
    g = load_golden("fpn")
    net = FeatureNetwork(_conf())
    assert sorted(net.state_dict().keys()) == sorted(g["sd"].keys()), "parameter names must equal the reference's"
    net.load_state_dict(g["sd"], strict=True)
    net = net.to(DEV)
    outs = net(g["in"]["imgs"].to(DEV))
    assert len(outs) == 4
    for i, o in enumerate(outs):
        want = g["out"]["feat%d" % i]
        assert tuple(o.shape) == want.shape
        assert_close(o, want, 1e-4, "FeatureNetwork stage %d (coarse -> fine) vs the reference" % i)
    # bitwise reproducible (the statistics are summed in a fixed order)
    again = net(g["in"]["imgs"].to(DEV))
    assert all(torch.equal(a, b) for a, b in zip(outs, again))


@pytest.mark.parametrize("nv,H,W", [(3, 64, 88), (1, 16, 24), (5, 120, 160)])
def test_other_sizes_vs_the_torch_restatement(nv, H, W):
    torch.manual_seed(H)
    net = FeatureNetwork(_conf())
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.rand(nv, 3, H, W)
    want = fpn_oracle.feature_network_forward(sd, x)
    outs = net.to(DEV)(x.to(DEV))
    for i, (o, w) in enumerate(zip(outs, want)):
        assert_close(o, w, 2e-4, "stage %d at %dx%d" % (i, H, W))


def test_full_size_images_and_errors():
    net = FeatureNetwork(_conf()).to(DEV)
    x = torch.rand(3, 3, 576, 800, device=DEV)
    outs = net(x)
    assert [tuple(o.shape) for o in outs] == [(3, 4, 72, 100), (3, 4, 144, 200), (3, 4, 288, 400), (3, 4, 576, 800)]
    assert all(bool(torch.isfinite(o).all()) for o in outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net(x)
    e1.record()
    torch.cuda.synchronize()
    print("FeatureNetwork 3 x 576 x 800: %.2f ms" % (e0.elapsed_time(e1) / 5))
    with pytest.raises(ValueError):
        net(torch.rand(1, 3, 50, 64, device=DEV))
    with pytest.raises(RuntimeError):
        net(torch.rand(1, 3, 64, 64))


def test_surf_model_owns_the_feature_network():
    """SuRF built from a reference conf has `feature_network` under the reference's attribute name (surf.py:25), so the
    `feature_network.*` keys of a reference checkpoint load; extract_features = surf.py:69."""
    import os
    from surf_b200.surf import SuRF
    here = os.path.dirname(os.path.abspath(__file__))
    c = conf.parse_file(os.path.join(here, "..", "surf_b200", "confs", "surf.conf")) if os.path.exists(
        os.path.join(here, "..", "surf_b200", "confs", "surf.conf")) else None
    if c is None:
        c = conf.ConfigTree()
        m = conf.ConfigTree()
        m.put("range_ratios", [1.0, 0.4, 0.1, 0.01])
        m.put("implicit_surface", conf.default_implicit_surface_conf())
        m.put("feature_network", _conf())
        c.put("model", m)
    g = load_golden("fpn")
    model = SuRF(c["model"])
    model.feature_network.load_state_dict(g["sd"], strict=True)
    assert {"feature_network." + k for k in g["sd"]} <= set(model.state_dict().keys())
    model = model.to(DEV)
    outs = model.extract_features(g["in"]["imgs"].to(DEV))
    for i, o in enumerate(outs):
        assert_close(o, g["out"]["feat%d" % i], 1e-4, "SuRF.extract_features stage %d" % i)
