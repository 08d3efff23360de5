"""GPU marching cubes (replaces the host call ``mcubes.marching_cubes(u, threshold)`` of extract_geometry,
implicit_surface.py:353) and a minimal binary PLY writer for Runner.validate's mesh export (runner.py:236-243 uses
trimesh, which this image does not have)."""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def marching_cubes_device(u: torch.Tensor, threshold: float = 0.0, x_offset: int = 0):
    """u (nx,ny,nz) fp32 CUDA tensor -> (vertices (nv,3) float64, triangles (nt,3) int32), device tensors.
    Vertices are in grid-index coordinates (PyMCubes' convention); a corner is inside when u > threshold."""
    if not u.is_cuda:
        raise RuntimeError("surf_b200.mesh.marching_cubes_device needs a CUDA tensor (there is no CPU fallback)")
    u = u.detach().to(torch.float32).contiguous()
    assert u.dim() == 3
    nx, ny, nz = (int(v) for v in u.shape)
    lib = _lib.load()
    with torch.cuda.device(u.device):
        ws_bytes = int(lib.surf_mc_workspace_bytes(nx, ny, nz))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=u.device)
        counts = torch.zeros(2, dtype=torch.int64, device=u.device)
        _lib.check(lib.surf_mc_count(u.data_ptr(), nx, ny, nz, float(threshold), ws.data_ptr(), ws_bytes,
                                     counts.data_ptr(), _stream()), "mc_count")
        nv, nt = (int(v) for v in counts.cpu())          # the one host sync: output sizes
        verts = torch.empty((nv, 3), dtype=torch.float64, device=u.device)
        tris = torch.empty((nt, 3), dtype=torch.int32, device=u.device)
        _lib.check(lib.surf_mc_emit(u.data_ptr(), nx, ny, nz, float(threshold), ws.data_ptr(), int(x_offset),
                                    verts.data_ptr(), nv, tris.data_ptr(), nt, _stream()), "mc_emit")
    return verts, tris


def marching_cubes(u, threshold: float = 0.0):
    """Drop-in for ``mcubes.marching_cubes``: numpy (or tensor) grid in, numpy ``(vertices float64, triangles int64)``
    out.  The grid is uploaded when it is not already on the device."""
    if not isinstance(u, torch.Tensor):
        u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float32))
    if not u.is_cuda:
        u = u.cuda()
    v, t = marching_cubes_device(u, threshold)
    return v.cpu().numpy(), t.cpu().numpy().astype(np.int64)


def write_ply(path: str, vertices: np.ndarray, triangles: np.ndarray) -> None:
    """Binary little-endian PLY (float32 x/y/z + triangle faces), what ``trimesh.Trimesh(v, f).export('x.ply')``
    produces for a bare mesh."""
    v = np.ascontiguousarray(vertices, dtype="<f4")
    f = np.ascontiguousarray(triangles, dtype="<i4")
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
           "property float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(v), len(f)))
    rec = np.empty(len(f), dtype=[("n", "u1"), ("i", "<i4", (3,))])
    rec["n"] = 3
    rec["i"] = f
    with open(path, "wb") as fh:
        fh.write(hdr.encode("ascii"))
        fh.write(v.tobytes())
        fh.write(rec.tobytes())
