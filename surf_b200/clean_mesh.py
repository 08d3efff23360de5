"""GPU edition of the reference's mesh cleaning (utils/clean_mesh.py:10-129; Runner.validate calls it with
``--clean_mesh``, runner.py:233).  Same three stages, same parameters:

    dilate the object masks with a disk                          surf_mask_dilate
    clean_mesh_by_mask        (vertex visible in > n views)      surf_mesh_vertex_visibility
    clean_mesh_outside_frustum (first hit of a masked ray,       surf_mesh_first_hits  (z-buffer instead of embree)
                                components of >= 500 faces)      surf_mesh_components  (edge hash + union-find)

The reference needs skimage, trimesh (+ pyembree) and open3d; none of them is required here.  torch is used for buffer
management and boolean-mask compaction only."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _host_f32(x):
    return np.ascontiguousarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x, dtype=np.float32)


def dilate_masks(masks: torch.Tensor, radius: int) -> torch.Tensor:
    """(nv,H,W) bool/byte CUDA tensor -> dilated byte tensor (skimage.morphology.binary_dilation(m, disk(radius)))."""
    m = (masks != 0).to(torch.uint8).contiguous()
    nv, h, w = (int(v) for v in m.shape)
    out = torch.empty_like(m)
    ws = torch.empty((nv, h, w), dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        _lib.check(_lib.load().surf_mask_dilate(m.data_ptr(), nv, h, w, int(radius), ws.data_ptr(), out.data_ptr(), _stream()),
                   "mask_dilate")
    return out


def vertex_visibility(vertices: torch.Tensor, masks: torch.Tensor, intrs, c2ws) -> torch.Tensor:
    """Per vertex: the number of views in which it projects inside the image and onto the mask (clean_mesh.py:12-28)."""
    v = vertices.to(torch.float32).contiguous()
    m = (masks != 0).to(torch.uint8).contiguous()
    nv, h, w = (int(x) for x in m.shape)
    c2w = torch.as_tensor(_host_f32(c2ws))
    w2c = np.ascontiguousarray(torch.linalg.inv(c2w)[:, :3, :].numpy(), dtype=np.float32)      # (nv,3,4) host
    K = np.ascontiguousarray(_host_f32(intrs)[:, :3, :3])
    count = torch.empty(v.shape[0], dtype=torch.int32, device=v.device)
    if v.shape[0] == 0:
        return count
    with torch.cuda.device(v.device):
        _lib.check(_lib.load().surf_mesh_vertex_visibility(v.data_ptr(), v.shape[0], w2c.ctypes.data, K.ctypes.data, nv,
                                                           m.data_ptr(), h, w, count.data_ptr(), _stream()),
                   "mesh_vertex_visibility")
    return count


def first_hit_faces(vertices: torch.Tensor, faces: torch.Tensor, masks: torch.Tensor, intrs, c2ws, upscale=2):
    """(face_hit (F,) bool, n_missed_rays): faces that are the first hit of a masked camera ray of any view
    (clean_mesh.py:41-78)."""
    lib = _lib.load()
    v = vertices.to(torch.float32).contiguous()
    f = faces.to(torch.int32).contiguous()
    m = (masks > 0).to(torch.uint8).contiguous()
    nv, h, w = (int(x) for x in m.shape)
    hs, ws_ = int(h * upscale), int(w * upscale)
    c2w = torch.as_tensor(_host_f32(c2ws))
    w2c = np.ascontiguousarray(torch.linalg.inv(c2w)[:, :3, :].numpy(), dtype=np.float32)
    c2w_h = np.ascontiguousarray(c2w[:, :3, :].numpy(), dtype=np.float32)
    K = np.ascontiguousarray(_host_f32(intrs)[:, :3, :3])
    hit = torch.zeros(max(1, f.shape[0]), dtype=torch.uint8, device=v.device)
    stats = torch.zeros(2, dtype=torch.int32, device=v.device)
    if f.shape[0] == 0:                  # nothing to hit: every masked ray misses
        return hit[:0].bool(), int((m > 0).sum()) * int(upscale) ** 2
    with torch.cuda.device(v.device):
        nbytes = int(lib.surf_mesh_raster_workspace_bytes(hs, ws_))
        wsb = torch.empty(nbytes, dtype=torch.uint8, device=v.device)
        for i in range(nv):
            _lib.check(lib.surf_mesh_first_hits(v.data_ptr(), f.data_ptr(), f.shape[0], w2c[i].ctypes.data,
                                                c2w_h[i].ctypes.data, K[i].ctypes.data, m[i].data_ptr(), h, w, hs, ws_,
                                                wsb.data_ptr(), nbytes, hit.data_ptr(), stats.data_ptr(), _stream()),
                       "mesh_first_hits")
            if int(stats[1]) > 65536:
                raise RuntimeError("clean_mesh: more than 65536 faces cover over 4096 ray samples each (not a surface mesh?)")
    return hit[:f.shape[0]].bool(), int(stats[0])


def face_components(faces: torch.Tensor, min_len=500):
    """(label (F,) int32, keep (F,) bool) of the face-adjacency graph (trimesh.graph.connected_components on
    mesh.face_adjacency, clean_mesh.py:99)."""
    lib = _lib.load()
    f = faces.to(torch.int32).contiguous()
    n = int(f.shape[0])
    label = torch.empty(max(1, n), dtype=torch.int32, device=f.device)
    keep = torch.zeros(max(1, n), dtype=torch.uint8, device=f.device)
    if n:
        with torch.cuda.device(f.device):
            nbytes = int(lib.surf_mesh_components_workspace_bytes(n))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=f.device)
            _lib.check(lib.surf_mesh_components(f.data_ptr(), n, int(min_len), ws.data_ptr(), nbytes, label.data_ptr(),
                                                keep.data_ptr(), _stream()), "mesh_components")
    return label[:n], keep[:n].bool()


def clean_mesh(vertices, triangles, masks, intrs, c2ws, dilation_radius=11, min_nb_visible=1, upscale=2, min_len=500,
               device="cuda", return_stages=False):
    """utils/clean_mesh.py:110-129.  vertices (V,3), triangles (F,3) numpy arrays or tensors (e.g. from extract_geometry),
    masks (nv,H,W) or (nv,H,W,C).  Returns numpy (vertices float64, triangles int64) of the cleaned mesh with the
    unreferenced vertices removed."""
    if not torch.cuda.is_available():
        raise RuntimeError("surf_b200.clean_mesh needs a CUDA device (there is no CPU fallback)")
    dev = torch.device(device)
    v64 = torch.as_tensor(np.asarray(vertices) if not isinstance(vertices, torch.Tensor) else vertices).to(dev)
    v = v64.to(torch.float32)                  # torch.from_numpy(mesh.vertices).float() (:12)
    f = torch.as_tensor(np.asarray(triangles) if not isinstance(triangles, torch.Tensor) else triangles).to(dev).to(torch.int64)
    m = torch.as_tensor(masks).to(dev).float()
    if m.dim() > 3:
        m = m.mean(dim=-1)                     # :116-117
    stages = {}
    # stage 1+2: vertices seen inside the dilated masks of more than min_nb_visible views (:119-125, :10-34)
    dil = dilate_masks(m > 0.5, dilation_radius)
    count = vertex_visibility(v, dil, intrs, c2ws)
    valid = count > int(min_nb_visible)
    f = f[valid[f].all(dim=-1)]
    stages["faces_after_mask"] = int(f.shape[0])
    # stage 3: first hits of the masked camera rays (:38-96); the un-dilated masks, `mask > 0`
    hit, missed = first_hit_faces(v, f, m, intrs, c2ws, upscale)
    if missed == 0 and bool(hit.any()):
        # the reference drops the smallest of the sorted hit values (`values[1:]`, :94), which is the miss marker -1 —
        # or, when every ray hit the mesh, the face with the smallest index
        hit[int(torch.nonzero(hit)[0])] = False
    f = f[hit]
    stages["faces_after_frustum"] = int(f.shape[0])
    # stage 4: connected components of at least min_len faces (:99-102), then remove_unreferenced_vertices (:103)
    _, keep = face_components(f, min_len)
    f = f[keep]
    stages["faces_after_components"] = int(f.shape[0])
    used = torch.zeros(v.shape[0], dtype=torch.bool, device=dev)
    used[f.reshape(-1)] = True
    remap = torch.cumsum(used.to(torch.int64), 0) - 1
    out_v = v64[used].double().cpu().numpy()
    out_f = remap[f].cpu().numpy().astype(np.int64)
    return (out_v, out_f, stages) if return_stages else (out_v, out_f)
