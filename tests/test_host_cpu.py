"""CPU-side checks: conf reader, synthetic scenes, module parameter names, C-ABI exports, RNG stream."""
import ctypes
import os
import re

import pytest
import torch

from helpers import GOLDEN, load_golden
from surf_b200 import _lib, conf, synthetic
from surf_b200.modules.implicit_surface import ImplicitSurface

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_conf_default_block():
    c = conf.default_implicit_surface_conf()
    assert c.get_list("render.n_samples") == [64, 32, 24, 16]
    assert c.get_float("render.perturb") == 1.0
    assert dict(**c["color_network"]) == {"d_feature": 16}
    assert c["sdf_network"]["skip_in"] == [3]
    assert c.get_bool("missing", default=False) is False
    with pytest.raises(conf.ConfigMissingException):
        c.get_int("nope")


def test_conf_syntax_subset():
    t = conf.parse_string("""
    a { b = 1  # comment
        c = [1, 2.5, x]
        d{ e = True }
    }
    s = <some path>
    q : "quoted # not comment"
    """)
    assert t["a.b"] == 1 and t["a.c"] == [1, 2.5, "x"] and t["a.d.e"] is True
    assert t.get_string("s") == "<some path>" and t["q"] == "quoted # not comment"
    assert "a.d" in t and "a.z" not in t


def test_synthetic_scene_layout():
    sc = synthetic.make_scene(3, 48, 64, 8, seed=1)
    assert [v.shape[1] for v in sc.volumes] == [7] * 4
    dims = [i.shape[0] for i in sc.sparse_idxes]
    assert dims == [64, 32, 16, 8]                      # fine -> coarse
    for v, i, m in zip(sc.volumes, sc.sparse_idxes, sc.mask_volumes):
        assert i.dtype == torch.int64 and m.shape == (1, 1) + tuple(i.shape)
        occ = i >= 0
        assert int(occ.sum()) == v.shape[0]
        assert torch.equal(occ, m[0, 0] > 0)
        assert torch.equal(i[occ], torch.arange(v.shape[0]))
    assert [tuple(f.shape[-2:]) for f in sc.features] == [(48, 64), (24, 32), (12, 16), (6, 8)]
    o, d = synthetic.random_pixel_rays(sc, 5)
    assert torch.allclose(d.norm(dim=-1), torch.ones(5), atol=1e-6)


def test_module_loads_reference_state_dict():
    g = load_golden("render_v2_perturbed")
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    assert set(m.state_dict().keys()) == set(g["sd"].keys())
    m.load_state_dict(g["sd"], strict=True)
    for k, v in m.state_dict().items():
        assert v.shape == g["sd"][k].shape, k


def test_chunk_random_stream_matches_reference_order():
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    torch.manual_seed(3)
    a = m.draw_chunk_randoms(600, 256)
    torch.manual_seed(3)
    rows = []
    for b in (256, 256, 88):
        rows.append(torch.cat([torch.rand([b, 1]) for _ in range(4)], 1))
        torch.rand([1024, 3])
    assert torch.equal(a, torch.cat(rows))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "surf_b200.h")).read()
    declared = set(re.findall(r"\b(surf_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -m surf_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().surf_version() == _lib.ABI_VERSION == 4


def test_struct_sizes_match_header():
    # mirrors of the POD structs: a size drift means the ctypes binding no longer matches the header
    assert ctypes.sizeof(_lib.RenderOutputs) == 22 * 8
    assert ctypes.sizeof(_lib.DepthMapParams) == 4 * (9 + 9 + 3 + 3 + 2 + 2 + 6)
    assert ctypes.sizeof(_lib.ExtrasParams) == 4 * (9 + 3 + 9 + 9 + 8 * 9 + 8 * 9 + 8 * 3 + 2)
    assert ctypes.sizeof(_lib.RenderCfg) == 4 + 16 + 16 + 4 + 4 + 4 + 4 + 4 + 8 + 8  # incl. alignment padding
    assert ctypes.sizeof(_lib.SceneStats) == 9 * 8


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    with pytest.raises(RuntimeError):
        m.net_handle()
