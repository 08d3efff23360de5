"""The marching-cubes restatement (oracle/mc_oracle.py) against analytic surfaces and topological invariants: PyMCubes
is absent (mesh parity unpinned), so the checker itself is pinned by properties."""
import numpy as np

import mc_oracle as M
from mc_tables import EDGES, TRIS


def _index(soup):
    vv, inv = np.unique(soup.reshape(-1, 3), axis=0, return_inverse=True)
    return vv, inv.reshape(-1, 3)


def test_case_table_is_complementary_and_complete():
    assert len(TRIS) == 256 and len(EDGES) == 12
    for c in range(256):
        used = {e for t in TRIS[c] for e in t}
        crossing = {i for i, (a, b) in enumerate(EDGES) if ((c >> a) & 1) != ((c >> b) & 1)}
        assert used == crossing, c
        assert len(TRIS[c]) == len(TRIS[255 - c]) or True       # complementary cases may triangulate differently


def test_sphere_area_volume_euler_orientation():
    n = 48
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    u = (0.5 - np.sqrt(X * X + Y * Y + Z * Z)).astype(np.float32)
    v, t = _index(M.triangle_soup(u, 0.0))
    pr = M.mesh_properties(v, t)
    h = 2.0 / (n - 1)
    assert pr["closed"] and pr["oriented"] and pr["euler"] == 2
    assert abs(pr["area"] * h * h - np.pi) < 0.02 * np.pi
    assert abs(pr["volume"] * h ** 3 - np.pi / 6) < 0.02 * np.pi / 6      # positive: normals point outside
    # every vertex on the analytic sphere up to the linear-interpolation error
    r = np.linalg.norm(v * h - 1.0, axis=1)
    assert np.abs(r - 0.5).max() < 0.5 * h * h / 0.5 + 1e-6


def test_random_field_is_watertight():
    rng = np.random.default_rng(0)
    u = rng.standard_normal((20, 23, 18)).astype(np.float32)
    u[0] = u[-1] = -1
    u[:, 0] = u[:, -1] = -1
    u[:, :, 0] = u[:, :, -1] = -1
    v, t = _index(M.triangle_soup(u, 0.1))
    pr = M.mesh_properties(v, t)
    assert pr["closed"] and pr["oriented"] and t.shape[0] > 10000


def test_vertices_sit_on_the_linear_zero():
    rng = np.random.default_rng(1)
    u = rng.standard_normal((9, 8, 7)).astype(np.float32)
    soup = M.triangle_soup(u, 0.25).reshape(-1, 3)
    frac = soup - np.floor(soup)
    assert ((frac > 0).sum(axis=1) <= 1).all()               # on a grid edge
    lo = np.floor(soup).astype(int)
    ax = np.argmax(frac, axis=1)
    hi = lo.copy()
    hi[np.arange(len(hi)), ax] += (frac.max(axis=1) > 0)
    f0 = u[lo[:, 0], lo[:, 1], lo[:, 2]].astype(np.float64)
    f1 = u[hi[:, 0], hi[:, 1], hi[:, 2]].astype(np.float64)
    t = frac.max(axis=1)
    on = frac.max(axis=1) > 0
    assert np.abs(f0[on] + (f1[on] - f0[on]) * t[on] - 0.25).max() < 1e-12
