"""GPU mesh cleaning (csrc/mesh_clean.cu, surf_b200/clean_mesh.py; the reference: utils/clean_mesh.py:10-129) against the
CPU restatement oracle/clean_mesh_oracle.py, stage by stage and end to end.  Integer / index results: equal, or every
difference shown to sit on a decision boundary of the reference's float arithmetic (census with proof)."""
import ctypes as C

import numpy as np
import pytest
import torch

import clean_mesh_oracle as CO
from surf_b200 import _lib, clean_mesh as CM, mesh, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _blob_mesh(res=72):
    """One large blob + four small far-away ones (each its own component), meshed by the GPU marching cubes; world units."""
    ax = torch.linspace(-1, 1, res, device=DEV)
    x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
    u = 0.55 - torch.sqrt(x ** 2 + (y * 1.2) ** 2 + z ** 2) + 0.05 * torch.sin(9 * x) * torch.sin(7 * y)
    for cx, cy, cz in ((0.8, 0.8, 0.0), (-0.8, 0.75, 0.1), (0.78, -0.8, -0.1), (-0.8, -0.8, 0.0)):
        u = torch.maximum(u, 0.07 - torch.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2))
    v, t = mesh.marching_cubes_device(u.contiguous(), 0.0)
    v = v / (res - 1) * 2.0 - 1.0
    return v.cpu().numpy(), t.cpu().numpy().astype(np.int64)


def _views(nv=4, H=60, W=80):
    intrs, c2ws, _, _ = synthetic.make_cameras(nv, H, W)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    masks = []
    for i in range(nv):        # an ellipse that cuts the silhouette of the big blob, shifted per view; one hole
        m = (((xx - W / 2 - 3 * i) / (0.33 * W)) ** 2 + ((yy - H / 2 + 2 * i) / (0.42 * H)) ** 2) <= 1.0
        m &= ~(((xx - W / 2) ** 2 + (yy - H / 2 - 6) ** 2) <= 9)
        masks.append(m)
    return torch.stack(masks).float(), intrs, c2ws


@pytest.mark.parametrize("radius", [0, 1, 3, 11])
def test_disk_dilation_equals_scipy(radius):
    g = torch.Generator().manual_seed(radius)
    m = torch.rand(3, 45, 70, generator=g) < 0.01
    m[0, 0, 0] = True
    m[1, 44, 69] = True
    m[2] = False
    got = CM.dilate_masks(m.to(DEV), radius).cpu().numpy().astype(bool)
    want = CO.dilate(m.numpy(), radius)
    assert np.array_equal(got, want)


def test_vertex_visibility_vs_oracle():
    v, t = _blob_mesh()
    masks, intrs, c2ws = _views()
    dil = torch.from_numpy(CO.dilate((masks > 0.5).numpy(), 4))
    count = CM.vertex_visibility(torch.from_numpy(v).to(DEV), dil.to(DEV), intrs, c2ws).cpu()
    valid, want, margin = CO.vertex_valid(v, dil, intrs, c2ws, 1, return_margin=True)
    mism = count != want
    assert int(want.max()) >= 3 and int(want.min()) == 0, "the scene should have visible and invisible vertices"
    # a count may only differ where a projection sits on an integer pixel coordinate (the bilinear footprint flips)
    assert bool((margin[mism] < 1e-3).all()), "unexplained visibility mismatches: %d" % int((margin[mism] >= 1e-3).sum())
    assert float(mism.float().mean()) < 1e-3
    print("vertex visibility: %d vertices, %d boundary mismatches" % (len(v), int(mism.sum())))


def test_first_hits_per_ray_vs_oracle():
    """The z-buffer of one view, ray by ray, against the brute-force ray caster."""
    lib = _lib.load()
    v, t = _blob_mesh(56)
    masks, intrs, c2ws = _views(2, 40, 52)
    up, i = 2, 1
    h, w = masks.shape[1:]
    hs, ws = h * up, w * up
    vd = torch.from_numpy(v).float().to(DEV).contiguous()
    fd = torch.from_numpy(t).int().to(DEV).contiguous()
    m = (masks[i] > 0).to(torch.uint8).to(DEV).contiguous()
    w2c = np.ascontiguousarray(torch.linalg.inv(c2ws)[i, :3, :].numpy(), dtype=np.float32)
    c2w = np.ascontiguousarray(c2ws[i, :3, :].numpy(), dtype=np.float32)
    K = np.ascontiguousarray(intrs[i, :3, :3].numpy(), dtype=np.float32)
    nbytes = int(lib.surf_mesh_raster_workspace_bytes(hs, ws))
    wsb = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    hit = torch.zeros(len(t), dtype=torch.uint8, device=DEV)
    stats = torch.zeros(2, dtype=torch.int32, device=DEV)
    _lib.check(lib.surf_mesh_first_hits(vd.data_ptr(), fd.data_ptr(), len(t), w2c.ctypes.data, c2w.ctypes.data,
                                        K.ctypes.data, m.data_ptr(), h, w, hs, ws, wsb.data_ptr(), nbytes, hit.data_ptr(),
                                        stats.data_ptr(), torch.cuda.current_stream().cuda_stream), "first_hits")
    z = wsb[:hs * ws * 8].view(torch.int64).cpu().numpy().view(np.uint64)
    got = np.where(z == np.uint64(0xffffffffffffffff), -1, (z & np.uint64(0xffffffff)).astype(np.int64))
    ro, rd = CO.camera_rays(intrs[i], c2ws[i], h, w, up)
    sel = (torch.nn.functional.interpolate(masks[i][None, None], scale_factor=up, mode="nearest")[0, 0] > 0).view(-1).numpy()
    want, margin = CO.first_hits(torch.from_numpy(v).float().numpy(), t, ro.numpy(), rd.numpy())
    assert (got[~sel] == -1).all(), "rays outside the mask must not be cast"
    mism = sel & (got != want)
    assert int((want[sel] >= 0).sum()) > 500 and int((want[sel] < 0).sum()) > 50, "the view should have hits and misses"
    assert (margin[mism] < 1e-4).all(), "unexplained first-hit mismatches: %s" % np.nonzero(mism & (margin >= 1e-4))[0][:10]
    assert mism.mean() < 2e-3
    # the face flags and the miss counter follow from the z-buffer
    flags = np.zeros(len(t), dtype=bool)
    flags[got[got >= 0]] = True
    assert np.array_equal(hit.cpu().numpy().astype(bool), flags)
    assert int(stats[0]) == int((sel & (got < 0)).sum())
    print("first hits: %d rays, %d boundary mismatches" % (int(sel.sum()), int(mism.sum())))


def test_large_faces_and_faces_behind_the_camera():
    """A quad that fills the screen (footprint above the per-thread limit) in front of a small triangle, and a triangle
    that straddles the camera plane."""
    masks, intrs, c2ws = _views(1, 48, 64)
    masks[:] = 1.0
    v = np.array([[-3, -3, 0.5], [3, -3, 0.5], [3, 3, 0.5], [-3, 3, 0.5],           # big quad at z = 0.5
                  [-0.1, -0.1, 0.0], [0.1, -0.1, 0.0], [0.0, 0.1, 0.0],               # small triangle in front of it
                  [0.2, 0.2, -1.0], [0.3, 0.2, -3.0], [0.2, 0.3, -3.0]], dtype=np.float64)    # crosses the camera plane
    t = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.int64)
    hit, missed = CM.first_hit_faces(torch.from_numpy(v).to(DEV), torch.from_numpy(t).to(DEV), masks.to(DEV), intrs, c2ws, 2)
    ro, rd = CO.camera_rays(intrs[0], c2ws[0], 48, 64, 2)
    want, _ = CO.first_hits(v.astype(np.float32), t, ro.numpy(), rd.numpy())
    flags = np.zeros(4, dtype=bool)
    flags[want[want >= 0]] = True
    assert flags[:3].all(), "the test geometry should be hit"
    assert np.array_equal(hit.cpu().numpy(), flags)
    assert missed == int((want < 0).sum())


def test_components_vs_scipy():
    v, t = _blob_mesh()
    rng = np.random.default_rng(0)
    t = t[rng.random(len(t)) > 0.02]                     # punch holes: some faces lose neighbours
    t = np.concatenate([t, t[:3]])                       # duplicated faces: edges shared by more than two faces
    label, keep = CM.face_components(torch.from_numpy(t).to(DEV), 500)
    want_keep, want_lab = CO.components_keep(t, 500)
    label, keep = label.cpu().numpy(), keep.cpu().numpy()
    assert np.array_equal(keep, want_keep)
    assert 0 < keep.sum() < len(t)
    # same partition of the graph nodes (labels up to renaming)
    adj = CO.face_adjacency(t)
    nodes = np.unique(adj)
    pairs = set(zip(label[nodes].tolist(), want_lab[nodes].tolist()))
    assert len(pairs) == len(set(label[nodes].tolist())) == len(set(want_lab[nodes].tolist()))
    assert (label[nodes] <= nodes).all()                 # the label is the smallest face index of the component


@pytest.mark.parametrize("min_visible", [1, 0])
def test_clean_mesh_end_to_end(min_visible):
    v, t = _blob_mesh(60)
    masks, intrs, c2ws = _views(3, 48, 64)
    gv, gt, st = CM.clean_mesh(v, t, masks, intrs, c2ws, dilation_radius=3, min_nb_visible=min_visible, upscale=2,
                               min_len=500, return_stages=True)
    wv, wt = CO.clean_mesh(v, t, masks, intrs, c2ws, dilation_radius=3, min_nb_visible=min_visible, upscale=2, min_len=500)
    print("clean_mesh:", len(t), "faces ->", st, "oracle:", len(wt))
    assert 500 <= len(wt) < len(t), "the test scene should remove some faces and keep the large component"
    assert gv.dtype == np.float64 and gt.dtype == np.int64 and gt.max() < len(gv) and len(np.unique(gt)) == len(gv)
    a = {tuple(map(tuple, np.round(gv[f], 9))) for f in gt}
    b = {tuple(map(tuple, np.round(wv[f], 9))) for f in wt}
    # the stages are compared ray by ray / vertex by vertex above; end to end the few boundary decisions may move
    # single faces in or out
    assert len(a ^ b) <= max(2, len(b) // 500), "face sets differ by %d of %d" % (len(a ^ b), len(b))
    # the small far-away blobs are gone, either by the mask or by the component filter
    assert float(np.abs(gv).max()) < 0.75


def test_clean_mesh_input_types():
    v, t = _blob_mesh(40)
    masks, intrs, c2ws = _views(3, 30, 40)
    m4 = masks[..., None].repeat(1, 1, 1, 3)             # (nv,H,W,C) masks are averaged over C (clean_mesh.py:116)
    a = CM.clean_mesh(torch.from_numpy(v), torch.from_numpy(t), m4, intrs, c2ws, 2, 1, 2, 100)
    b = CM.clean_mesh(v, t, masks, intrs, c2ws, 2, 1, 2, 100)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_degenerate_inputs():
    """Empty meshes, meshes that lose every face, masks that select nothing."""
    masks, intrs, c2ws = _views(2, 24, 32)
    v0, t0 = np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64)
    gv, gt = CM.clean_mesh(v0, t0, masks, intrs, c2ws, 2, 1, 2, 10)
    assert gv.shape == (0, 3) and gt.shape == (0, 3)
    v, t = _blob_mesh(32)
    gv, gt = CM.clean_mesh(v, t, torch.zeros_like(masks), intrs, c2ws, 2, 1, 2, 10)       # nothing is visible anywhere
    assert gv.shape == (0, 3) and gt.shape == (0, 3)
    gv, gt = CM.clean_mesh(v, t, masks, intrs, c2ws, 2, 1, 2, 10 ** 9)                     # no component is large enough
    assert gv.shape == (0, 3) and gt.shape == (0, 3)
    label, keep = CM.face_components(torch.zeros((0, 3), dtype=torch.int64, device=DEV), 5)
    assert label.numel() == 0 and keep.numel() == 0
    # faces that share no edge are not nodes of the adjacency graph: never kept, whatever min_len
    iso = torch.tensor([[0, 1, 2], [3, 4, 5]], device=DEV)
    _, keep = CM.face_components(iso, 1)
    assert not bool(keep.any())
