"""GPU marching cubes (csrc/marching.cu, replaces mcubes.marching_cubes of implicit_surface.py:353) against the CPU
restatement: the same triangles bit for bit, plus the size-independent mesh properties at sizes the CPU cannot check."""
import numpy as np
import pytest
import torch

import mc_oracle as M
from helpers import load_golden, scene_from_recipe
from surf_b200 import mesh

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _soup(v, t):
    return v[t.reshape(-1)].reshape(-1, 3, 3)


@pytest.mark.parametrize("shape,thr", [((20, 23, 18), 0.1), ((33, 17, 40), -0.3), ((2, 2, 2), 0.0), ((5, 1, 7), 0.0),
                                       ((64, 64, 65), 0.0)])
def test_same_triangles_as_the_oracle(shape, thr):
    rng = np.random.default_rng(sum(shape))
    u = rng.standard_normal(shape).astype(np.float32)
    v, t = mesh.marching_cubes(u, thr)
    ref = M.triangle_soup(u, thr)
    assert t.shape[0] == ref.shape[0], "triangle count"
    assert v.dtype == np.float64 and t.dtype == np.int64
    if t.shape[0]:
        assert int(t.max()) < v.shape[0] and int(t.min()) >= 0
        assert np.array_equal(M.canonical(_soup(v, t)), M.canonical(ref)), "triangle sets differ"
        assert np.unique(v, axis=0).shape[0] == v.shape[0], "vertices must be shared, not duplicated"
        assert np.unique(t).shape[0] == v.shape[0], "every vertex is referenced"


def test_empty_and_full_grids():
    for val in (-1.0, 1.0):
        v, t = mesh.marching_cubes(np.full((16, 16, 16), val, dtype=np.float32), 0.0)
        assert v.shape == (0, 3) and t.shape == (0, 3)


def test_values_equal_to_the_threshold_are_outside():
    u = np.zeros((8, 8, 8), dtype=np.float32)
    u[3:5, 3:5, 3:5] = 1.0
    v, t = mesh.marching_cubes(u, 0.0)
    pr = M.mesh_properties(v, t)
    assert pr["closed"] and pr["oriented"] and pr["euler"] == 2 and pr["volume"] > 0


def test_sphere_256_properties():
    n = 256
    g = torch.linspace(-1, 1, n, device=DEV)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    u = 0.5 - torch.sqrt(X * X + Y * Y + Z * Z)
    v, t = mesh.marching_cubes_device(u, 0.0)
    v, t = v.cpu().numpy(), t.cpu().numpy()
    pr = M.mesh_properties(v, t)
    h = 2.0 / (n - 1)
    assert pr["closed"] and pr["oriented"] and pr["euler"] == 2 and pr["all_vertices_used"]
    assert abs(pr["area"] * h * h - np.pi) < 2e-3 * np.pi
    assert abs(pr["volume"] * h ** 3 - np.pi / 6) < 2e-3 * np.pi / 6
    # deterministic
    v2, t2 = mesh.marching_cubes_device(u, 0.0)
    assert np.array_equal(v2.cpu().numpy(), v) and np.array_equal(t2.cpu().numpy(), t)


def test_x_slabs_with_a_halo_plane_tile_the_mesh():
    """The sharded form: slabs [x0, x1] that share one plane produce exactly the full grid's triangles."""
    rng = np.random.default_rng(5)
    u = torch.from_numpy(rng.standard_normal((40, 24, 24)).astype(np.float32)).to(DEV)
    vf, tf = mesh.marching_cubes_device(u, 0.0)
    full = M.canonical(_soup(vf.cpu().numpy(), tf.cpu().numpy().astype(np.int64)))
    parts = []
    for x0, x1 in ((0, 10), (10, 20), (20, 30), (30, 39)):
        v, t = mesh.marching_cubes_device(u[x0:x1 + 1], 0.0, x_offset=x0)
        parts.append(_soup(v.cpu().numpy(), t.cpu().numpy().astype(np.int64)))
    assert np.array_equal(M.canonical(np.concatenate(parts, axis=0)), full)


def test_extract_geometry_mesh_of_the_golden_scene():
    """extract_geometry end to end on the GPU: SDF grid (pinned against the reference golden elsewhere) -> mesh; the mesh
    equals the oracle's marching cubes of the same grid, is closed inside the volume and scaled to world units."""
    from surf_b200.modules.implicit_surface import ImplicitSurface
    from surf_b200 import conf
    g = load_golden("sdf_grid_24")
    sc = scene_from_recipe(g["recipe"])
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    m.load_state_dict(g["sd"])
    m = m.to(DEV)
    d = sc.to(DEV)
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    res = 48
    bmin, bmax = torch.tensor([-1.0, -1, -1]), torch.tensor([1.0, 1, 1])
    verts, tris = m.extract_geometry(ps, None, bmin, bmax, res, 0.0)
    u = m.sdf_grid(ps, bmin, bmax, res).cpu().numpy()
    ref = M.triangle_soup(u, 0.0)
    v_idx, t_idx = mesh.marching_cubes(u, 0.0)
    assert tris.shape[0] == ref.shape[0] and tris.shape[0] > 100
    assert np.array_equal(M.canonical(_soup(v_idx, t_idx)), M.canonical(ref))
    assert np.array_equal(t_idx, tris)
    np.testing.assert_allclose(verts, v_idx / (res - 1.0) * 2.0 - 1.0, rtol=0, atol=1e-12)     # implicit_surface.py:355
    assert np.abs(verts).max() <= 1.0
