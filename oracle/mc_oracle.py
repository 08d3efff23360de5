"""CPU restatement of marching cubes for the extract_geometry mesh — TEST INFRASTRUCTURE, never on the product path.

The reference calls ``mcubes.marching_cubes(u, threshold)`` (implicit_surface.py:353).  PyMCubes 0.1.4
(requirements.txt:154) is an un-vendored third-party dependency that is not installed here, so the mesh is
**parity unpinned** against the reference's own output; what is pinned:
  * the published algorithm (Lorensen & Cline 1987): one vertex per grid edge whose end points straddle the iso-value,
    placed at the linear zero  x1 + (iso - f1) / (f2 - f1)  (PyMCubes' ``mc_isovalue_interpolation``), triangles per cell
    from a 256-case table;
  * the case table (oracle/mc_tables.py) is generated from first principles by tools/gen_mc_tables.py with a face rule
    that makes neighbouring cells agree, hence a watertight surface;
  * size-independent properties checked in tests/: every mesh edge is shared by exactly two triangles with opposite
    orientation (closed, consistently oriented), Euler characteristic 2 and area / volume of an analytic sphere,
    normals pointing outside.
This numpy version loops over the cells in vectorised form and emits an unindexed triangle soup in fp64 — the GPU mesh
must consist of exactly the same triangles."""
import numpy as np

from mc_tables import EDGES, TRIS


def triangle_soup(u: np.ndarray, threshold: float = 0.0) -> np.ndarray:
    """(nx,ny,nz) -> (nt,3,3) fp64 triangle corner coordinates in grid-index units, cell order then table order."""
    u = np.asarray(u)
    nx, ny, nz = u.shape
    inside = u > np.float32(threshold)
    case = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.int32)
    for c in range(8):
        cx, cy, cz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= inside[cx:nx - 1 + cx, cy:ny - 1 + cy, cz:nz - 1 + cz].astype(np.int32) << c
    ud = u.astype(np.float64)
    thr = float(np.float32(threshold))
    out = []
    for cs in np.unique(case):
        tris = TRIS[int(cs)]
        if not tris:
            continue
        cells = np.argwhere(case == cs)                       # (m,3) lowest corner of each cell
        edge_pos = {}
        for e in {e for t in tris for e in t}:
            a, b = EDGES[e]
            pa = cells + np.array([a & 1, (a >> 1) & 1, (a >> 2) & 1])
            pb = cells + np.array([b & 1, (b >> 1) & 1, (b >> 2) & 1])
            fa = ud[pa[:, 0], pa[:, 1], pa[:, 2]]
            fb = ud[pb[:, 0], pb[:, 1], pb[:, 2]]
            t = (thr - fa) / (fb - fa)
            edge_pos[e] = pa.astype(np.float64) + (pb - pa) * t[:, None]
        for tri in tris:
            out.append((cells, np.stack([edge_pos[e] for e in tri], axis=1)))
    if not out:
        return np.zeros((0, 3, 3))
    return np.concatenate([t for _, t in out], axis=0)


def canonical(soup: np.ndarray) -> np.ndarray:
    """Order-independent form of a triangle soup: each triangle rotated so that its smallest corner comes first
    (orientation preserved), triangles sorted."""
    if soup.shape[0] == 0:
        return soup.reshape(0, 9)
    keys = soup[:, :, 0] * 1e12 + soup[:, :, 1] * 1e6 + soup[:, :, 2]
    first = np.argmin(keys, axis=1)
    idx = (first[:, None] + np.arange(3)[None, :]) % 3
    rot = np.take_along_axis(soup, idx[:, :, None], axis=1).reshape(-1, 9)
    order = np.lexsort(rot.T[::-1])
    return rot[order]


def mesh_properties(vertices: np.ndarray, triangles: np.ndarray):
    """-> dict(closed, oriented, euler, area, volume) of an indexed mesh."""
    t = np.asarray(triangles, dtype=np.int64)
    v = np.asarray(vertices, dtype=np.float64)
    he = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)          # directed half edges
    nvert = int(v.shape[0])
    key = he[:, 0] * nvert + he[:, 1]
    rkey = he[:, 1] * nvert + he[:, 0]
    uniq, cnt = np.unique(key, return_counts=True)
    oriented = bool((cnt == 1).all()) and bool(np.isin(rkey, uniq).all())            # each half edge once, twin present
    und = np.sort(he, axis=1)
    _, ucnt = np.unique(und[:, 0] * nvert + und[:, 1], return_counts=True)
    closed = bool((ucnt == 2).all())
    n_edges = int(ucnt.shape[0])
    used = np.unique(t)
    a, b, c = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    cr = np.cross(b - a, c - a)
    return {"closed": closed, "oriented": oriented, "euler": int(used.shape[0]) - n_edges + int(t.shape[0]),
            "area": float(0.5 * np.linalg.norm(cr, axis=1).sum()),
            "volume": float((a * cr).sum() / 6.0), "all_vertices_used": int(used.shape[0]) == nvert}
