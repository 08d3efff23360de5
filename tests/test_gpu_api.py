"""The reference-facing entry points of the drop-in: ImplicitSurface.forward in
val / train mode incl. the pseudo_pts branch, SuRF.forward after set_volumes (incl. view_ids), and the
Runner.validate-shaped consumer."""
import os

import numpy as np
import pytest
import torch

import surf_oracle as O
from helpers import RTOL_FP32, assert_close, load_golden, scene_from_recipe
from surf_b200 import _lib, conf, synthetic
from surf_b200.modules.implicit_surface import ImplicitSurface
from surf_b200.surf import SuRF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ipts(sc, rays_o, rays_d, **extra):
    d = {"imgs": sc.imgs, "intrs": sc.intrs, "c2ws": sc.c2ws, "rays_o": rays_o, "rays_d": rays_d, "near": sc.near,
         "far": sc.far}
    d.update(extra)
    return d


def _setup(name="render_v2_perturbed"):
    g = load_golden(name)
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"], strict=True)
    return g, sc, m.to(DEV), sc.to(DEV)


def test_forward_train_mode_with_pseudo_pts():
    """ImplicitSurface.forward(mode != 'val') = render() + pseudo_sdf of the masked pseudo points
    (implicit_surface.py:404-436)."""
    g, sc, m, d = _setup()
    i = g["in"]
    gen = torch.Generator().manual_seed(3)
    pseudo = torch.rand(2048, 3, generator=gen) * 2.6 - 1.3          # incl. points outside the volume
    pseudo[:4] = torch.tensor([[50.0, -80.0, 3.0], [1e4, 1e4, 1e4], [-1e6, 0.0, 0.0], [0.0, 0.0, 0.0]])
    ipts = _ipts(d, i["rays_o"].to(DEV), i["rays_d"].to(DEV), pseudo_pts=pseudo.to(DEV))
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = m("train", ipts, d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.features, d.features, 1.0, None)
    net = O.OracleNet(g["sd"])
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    ref = O.render(net, i["rays_o"], i["rays_d"], i["near"], i["far"], sc.matching_volume, sc.volumes, sc.sparse_idxes,
                   sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0)
    assert set(ref.keys()) <= set(out.keys())
    assert_close(out["gradients"], ref["gradients"], RTOL_FP32, "forward(train): gradients")
    assert_close(out["sparse_sdf"][:1024], ref["sparse_sdf"][:1024], RTOL_FP32, "forward(train): sparse_sdf (random points)")
    assert_close(out["color_fine"], ref["color_fine"], 5e-4, "forward(train): color_fine", floor=1e-2)
    # pseudo_sdf: sdf of the points inside the voxel mask, EXACTLY 0 elsewhere (:425-434) — also for far-away points
    pm = O.point_mask(pseudo, sc.mask_volumes)
    want = torch.zeros(pseudo.shape[0], 1)
    want[pm] = O.sdf_only(net, pseudo[pm], sc.volumes, sc.sparse_idxes)
    got = out["pseudo_sdf"].cpu()
    assert got.shape == (2048, 1)
    assert bool(torch.isfinite(got).all())
    assert bool((got[~pm] == 0).all()), "points outside the mask must be exactly 0"
    assert 0.05 < float(pm.float().mean()) < 0.95
    assert_close(got, want, RTOL_FP32, "pseudo_sdf")


def test_forward_val_mode_equals_validate():
    g = load_golden("validate_24x32")
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"], strict=True)
    m = m.to(DEV)
    d = sc.to(DEV)
    i = g["in"]
    r = int(g["recipe"]["res_level"])
    hw = (sc.H // r, sc.W // r)
    ipts = _ipts(d, i["rays_o"].to(DEV), i["rays_d"].to(DEV), bound_min=torch.tensor([-1.0, -1, -1]),
                 bound_max=torch.tensor([1.0, 1, 1]), hw=hw)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = m("val", ipts, d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.features, d.features)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    want = m.validate(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), d.matching_volume,
                      d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.features, d.intrs, d.c2ws,
                      ipts["bound_min"], ipts["bound_max"], hw, 1.0, None, extract_geometry=True, mesh_resolution=512)
    # the default val path extracts the mesh too (mesh_resolution 512 is hard-wired in the reference, :359)
    for k in ("img_fine", "normal_img", "sdf_depth", "render_depth"):
        assert isinstance(out[k], np.ndarray) and np.array_equal(out[k], want[k]), k
    assert torch.equal(out["color_fine"], want["color_fine"]) and out["color_fine"].device.type == "cpu"
    assert out["vertices"].ndim == 2 and out["vertices"].shape[1] == 3 and out["triangles"].shape[1] == 3
    assert out["vertices"].shape[0] > 1000, "the geometric-init sphere must produce a mesh"
    assert np.array_equal(out["triangles"], want["triangles"])


def _surf_model(g):
    c = conf.ConfigTree()
    c.put("range_ratios", [1.0, 0.4, 0.1, 0.01])
    c.put("implicit_surface", conf.default_implicit_surface_conf())
    model = SuRF(c)
    model.implicit_surface.load_state_dict(g["sd"], strict=True)
    return model.to(DEV)


def test_surf_forward_after_set_volumes_and_view_ids():
    """SuRF.forward dispatch (surf.py:133-163): lists arrive coarse->fine and are reversed for the renderer;
    `view_ids` selects feature maps per call without re-preparing the volume part of the scene."""
    g, sc, m, d = _setup()
    model = _surf_model(g)
    assert {k for k in model.state_dict() if k.startswith("implicit_surface.")} == {"implicit_surface." + k for k in g["sd"]}
    # build_volumes order: coarse -> fine (surf.py:80-131); features coarse -> fine
    model.set_volumes(d.volumes[::-1], d.sparse_idxes[::-1], d.mask_volumes[::-1], d.matching_volume, d.features[::-1])
    i = g["in"]
    ipts = _ipts(d, i["rays_o"].to(DEV), i["rays_d"].to(DEV))
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    out = model("train", ipts, 1.0, None)
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    want = m.render(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), d.matching_volume,
                    d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.features, d.intrs, d.c2ws, 1.0, None)
    for k in ("color_fine", "render_depth", "sdf_depth", "normal", "gradients", "weights"):
        assert torch.equal(out[k], want[k]), k
    # view_ids: identity selection -> same result; the volume part is not re-prepared (same scene object)
    from surf_b200.scene import GLOBAL_SCENE_CACHE
    before = dict(GLOBAL_SCENE_CACHE._d)
    ipts2 = dict(ipts, view_ids=torch.arange(sc.nv, device=DEV))
    for _ in range(3):
        torch.manual_seed(int(g["recipe"]["torch_seed"]))
        out2 = model("train", ipts2, 1.0, None)
    assert torch.equal(out2["color_fine"], want["color_fine"])
    after = dict(GLOBAL_SCENE_CACHE._d)
    assert set(before.keys()) == set(after.keys()) and all(before[k][2] is after[k][2] for k in before), \
        "new per-call feature tensors must not re-prepare the volume part"
    # a permutation of the source views permutes nothing in the output of a symmetric blend only approximately:
    # just check it runs and differs (different source order -> different reference view 0 stays)
    with pytest.raises(NotImplementedError):
        SuRF(model_conf()).forward("train", ipts)


def model_conf():
    c = conf.ConfigTree()
    c.put("implicit_surface", conf.default_implicit_surface_conf())
    return c


def test_sdf_network_forward_full_head():
    """SDFNetworkSparse.forward returns the reference's (n, 129) = [sdf / scale, 128 feature outputs]
    (sdf_network.py:95-121); column 0 is what sdf() returns."""
    g, sc, m, d = _setup()
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    pv = torch.from_numpy(g["out"]["_pts_valid"]).to(DEV)
    full = m.sdf_network(pv, ps)
    assert full.shape == (pv.shape[0], 129)
    assert_close(full, g["out"]["_sdf_full"], RTOL_FP32, "(n,129) head vs reference golden")
    assert_close(full[:, :1], m.sdf_network.sdf(pv, ps), 2e-5, "column 0 == sdf()")
    assert m.sdf_network(pv[:0], ps).shape == (0, 129)


def test_runner_validate_writes_reference_artifacts(tmp_path):
    """Runner.validate (runner.py:198-296): per scene a PNG image / normal map, depth maps (.npy + .png) and the
    mesh; here with the synthetic scene standing in for a dataset item."""
    from surf_b200.runner import validate
    g = load_golden("validate_24x32")
    sc = scene_from_recipe(g["recipe"])
    model = _surf_model(g)
    d = sc.to(DEV)
    model.set_volumes(d.volumes[::-1], d.sparse_idxes[::-1], d.mask_volumes[::-1], d.matching_volume, d.features[::-1])
    i = g["in"]
    r = int(g["recipe"]["res_level"])
    hw = (sc.H // r, sc.W // r)
    item = _ipts(d, i["rays_o"].to(DEV), i["rays_d"].to(DEV), bound_min=torch.tensor([-1.0, -1, -1]),
                 bound_max=torch.tensor([1.0, 1, 1]), hw=hw, file_name="scan1_0", scene="scan1",
                 scale_mat=torch.eye(4), color=torch.rand(hw[0] * hw[1], 3))
    torch.manual_seed(int(g["recipe"]["torch_seed"]))
    scalars = validate(model, [item], str(tmp_path), epoch=3, mesh_resolution=64)
    for sub, ext in (("val_img", "png"), ("val_normal", "png"), ("val_sdf_depth", "npy"), ("val_render_depth", "npy"),
                     ("val_sdf_depth", "png"), ("val_render_depth", "png")):
        assert os.path.exists(os.path.join(tmp_path, sub, "scan1_0_epoch3." + ext)), (sub, ext)
    assert os.path.exists(os.path.join(tmp_path, "meshes", "scan1_epoch3.ply"))
    rd = np.load(os.path.join(tmp_path, "val_render_depth", "scan1_0_epoch3.npy"))
    assert rd.shape == hw
    assert_close(rd, g["out"]["render_depth"], 2e-3, "render depth written by the runner vs the reference image", floor=1e-2)
    assert "psnr" in scalars and "color_loss" in scalars


def test_invalid_colour_path_is_rejected_and_empty_second_order_call():
    g, sc, m, d = _setup()
    i = g["in"]
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    m.color_path = 7
    with pytest.raises(RuntimeError):
        m.render(i["rays_o"].to(DEV), i["rays_d"].to(DEV), i["near"].to(DEV), i["far"].to(DEV), ps, None, None, None, None,
                 None, None, d.intrs, d.c2ws, 1.0, None)
    m.color_path = _lib.COLOR_SERIAL
    empty = torch.zeros((0, 3), device=DEV)
    assert m.sdf_network.smooth(empty, ps).shape == (0, 3)
    for mode in (_lib.MLP_TC, _lib.MLP_FFMA):
        gr, sm = m.sdf_network.smooth(empty, ps, with_grad=True, mode=mode)
        assert gr.shape == (0, 3) and sm.shape == (0, 3)
    with pytest.raises(RuntimeError):
        m.sdf_network.smooth(torch.zeros((4, 3), device=DEV), ps, mode=3)


def test_runner_validate_with_clean_mesh(tmp_path):
    """`--clean_mesh` (runner.py:233-234): the extracted mesh is cleaned against inputs["masks"] before it is written."""
    from surf_b200.runner import validate
    g = load_golden("validate_24x32")
    sc = scene_from_recipe(g["recipe"])
    model = _surf_model(g)
    d = sc.to(DEV)
    model.set_volumes(d.volumes[::-1], d.sparse_idxes[::-1], d.mask_volumes[::-1], d.matching_volume, d.features[::-1])
    i = g["in"]
    r = int(g["recipe"]["res_level"])
    hw = (sc.H // r, sc.W // r)
    yy, xx = torch.meshgrid(torch.arange(sc.H), torch.arange(sc.W), indexing="ij")
    masks = ((xx - sc.W / 2) ** 2 + (yy - sc.H / 2) ** 2 <= (0.3 * sc.H) ** 2).float()[None].repeat(sc.nv, 1, 1)

    def item():
        return _ipts(d, i["rays_o"].to(DEV), i["rays_d"].to(DEV), bound_min=torch.tensor([-1.0, -1, -1]),
                     bound_max=torch.tensor([1.0, 1, 1]), hw=hw, file_name="scan1_0", scene="scan1",
                     scale_mat=torch.eye(4), color=torch.rand(hw[0] * hw[1], 3), masks=masks)

    def n_faces(path):
        with open(path, "rb") as f:
            head = f.read(400).decode("latin1")
        return int(head.split("element face ")[1].split("\n")[0])

    torch.manual_seed(0)
    validate(model, [item()], str(tmp_path / "a"), epoch=1, mesh_resolution=64)
    torch.manual_seed(0)
    validate(model, [item()], str(tmp_path / "b"), epoch=1, mesh_resolution=64, clean_mesh=True)
    full = n_faces(os.path.join(tmp_path, "a", "meshes", "scan1_epoch1.ply"))
    cleaned = n_faces(os.path.join(tmp_path, "b", "meshes", "scan1_epoch1.ply"))
    assert 0 < cleaned < full, (cleaned, full)
