"""TEST INFRASTRUCTURE — CPU restatement of the reference's mesh cleaning (utils/clean_mesh.py:10-129), the checker of
surf_b200/clean_mesh.py.  Only tests/ may import it.

PARITY UNPINNED for the ray casting and the component filter: the reference delegates them to third-party code that is
absent from /root/reference and from this image — trimesh (`ray.ray_pyembree.RayMeshIntersector.intersects_first`,
`graph.connected_components` on `face_adjacency`), pyembree/embree2, skimage (`morphology.binary_dilation`, `disk`)
and open3d (imported, unused) — so utils/clean_mesh.py cannot even be imported here.  Their published algorithms are
restated: disk(r) = {dx^2 + dy^2 <= r^2} and scipy.ndimage.binary_dilation (what skimage calls); first hit = the
triangle with the smallest ray parameter t > 0 (Moller-Trumbore in float64, brute force over all faces); face adjacency =
pairs of faces sharing an edge that exactly two faces share, components through scipy.sparse.csgraph.  The torch part
(clean_mesh_by_mask, :10-34) is the reference's own sequence of torch calls."""
import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def disk(radius):                                   # skimage.morphology.disk
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return (X ** 2 + Y ** 2) <= radius ** 2


def dilate(masks_bool, radius):                     # clean_mesh.py:119-123
    return np.stack([ndimage.binary_dilation(m, structure=disk(radius)) for m in masks_bool])


def vertex_valid(vertices, masks, intrs, c2ws, min_nb_visible=1, return_margin=False):
    """clean_mesh_by_mask (:10-28): per-vertex validity; masks (nv,h,w) bool tensor."""
    points = torch.as_tensor(vertices).float().permute(1, 0)
    nv, h, w = masks.shape
    pts_cam = torch.matmul(c2ws.inverse(), torch.cat([points, torch.ones_like(points[:1])], dim=0)[None])[:, :3]
    pts_img = torch.matmul(intrs[:, :3, :3], pts_cam)
    pts_xy = pts_img[:, :2] / torch.clamp(pts_img[:, 2:], 1e-8)
    pix = pts_xy.clone()
    pts_xy[:, 0] = 2 * pts_xy[:, 0] / (w - 1) - 1
    pts_xy[:, 1] = 2 * pts_xy[:, 1] / (h - 1) - 1
    in_mask = (pts_xy.abs() <= 1).all(dim=1) & (pts_img[:, -1] > 1e-8)
    grid = torch.clamp(pts_xy.permute(0, 2, 1).unsqueeze(1), -10, 10)
    warp_mask = F.grid_sample(masks.unsqueeze(1).float(), grid, align_corners=True).squeeze(1).squeeze(1)
    count = ((warp_mask > 0) * in_mask).sum(dim=0)
    valid = count > min_nb_visible
    if return_margin:
        # distance (pixels) of every projection to the nearest integer pixel coordinate: where the bilinear footprint,
        # and with it `warp_mask > 0`, can flip under a rounding difference of the projection
        frac = (pix - pix.round()).abs().amin(dim=1)                       # (nv, np)
        return valid, count, frac.amin(dim=0)
    return valid


def camera_rays(intr, c2w, h, w, upscale):          # :47-64
    ys, xs = torch.meshgrid(torch.linspace(0, h - 1, int(h * upscale)), torch.linspace(0, w - 1, int(w * upscale)),
                            indexing="ij")
    p = torch.stack([xs, ys, torch.ones_like(ys)], dim=-1).view(-1, 3).float()
    p = torch.matmul(intr.inverse()[None, :3, :3], p[:, :, None]).squeeze(-1)
    rays_d = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    rays_d = torch.matmul(c2w[None, :3, :3], rays_d[:, :, None]).squeeze(-1)
    rays_o = c2w[None, :3, 3].expand(rays_d.shape)
    return rays_o, rays_d


def first_hits(vertices, faces, rays_o, rays_d, chunk=2048):
    """index of the first triangle along each ray (-1: none) + the margin of the decision: the smaller of the winning
    hit's distance to its triangle's border (barycentric units) and the relative gap to the runner-up hit."""
    V = np.asarray(vertices, dtype=np.float64)
    Fc = np.asarray(faces)
    a = V[Fc[:, 0]]
    e1 = V[Fc[:, 1]] - a
    e2 = V[Fc[:, 2]] - a
    o = np.asarray(rays_o, dtype=np.float64)
    d = np.asarray(rays_d, dtype=np.float64)
    idx = np.full(len(o), -1, dtype=np.int64)
    margin = np.full(len(o), np.inf)
    for s in range(0, len(o), chunk):
        oo, dd = o[s:s + chunk, None, :], d[s:s + chunk, None, :]
        p = np.cross(dd, e2[None])
        det = (e1[None] * p).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            t0 = oo - a[None]
            u = (t0 * p).sum(-1) * inv
            q = np.cross(t0, e1[None])
            v = (dd * q).sum(-1) * inv
            t = (e2[None] * q).sum(-1) * inv
        ok = (det != 0) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 0)
        tt = np.where(ok, t, np.inf)
        best = tt.argmin(axis=1)
        rows = np.arange(len(best))
        tb = tt[rows, best]
        hit = np.isfinite(tb)
        idx[s:s + chunk] = np.where(hit, best, -1)
        # margins: border distance of the winner, gap to the second-nearest hit, and for misses / all rays the
        # nearest "almost hit" (a triangle whose barycentric test fails by little)
        bd = np.minimum(np.minimum(u, v), 1 - u - v)                   # > 0 inside
        border = np.where(hit, bd[rows, best], np.inf)
        t2 = tt.copy()
        t2[rows, best] = np.inf
        second = t2.min(axis=1)
        with np.errstate(invalid="ignore"):
            gap = np.where(hit & np.isfinite(second), (second - tb) / np.maximum(tb, 1e-12), np.inf)
        near_miss = np.where((det != 0) & (t > 0) & ~ok, -bd, np.inf).min(axis=1)      # how far outside the closest non-hit is
        margin[s:s + chunk] = np.minimum(np.minimum(border, gap), near_miss)
    return idx, margin


def face_adjacency(faces):
    """trimesh.graph.face_adjacency: (n,2) pairs of faces sharing an edge that exactly two faces share."""
    Fc = np.asarray(faces)
    e = np.concatenate([Fc[:, [0, 1]], Fc[:, [1, 2]], Fc[:, [2, 0]]], axis=0)
    fid = np.tile(np.arange(len(Fc)), 3)
    ok = e[:, 0] != e[:, 1]
    e, fid = np.sort(e[ok], axis=1), fid[ok]
    key = e[:, 0].astype(np.int64) << 32 | e[:, 1].astype(np.int64)
    order = np.argsort(key, kind="stable")
    key, fid = key[order], fid[order]
    uniq, start, cnt = np.unique(key, return_index=True, return_counts=True)
    two = start[cnt == 2]
    return np.stack([fid[two], fid[two + 1]], axis=1) if len(two) else np.zeros((0, 2), dtype=np.int64)


def components_keep(faces, min_len=500):
    """mask of the faces in components (of the adjacency graph's nodes) with at least min_len faces (:99-102)."""
    n = len(faces)
    adj = face_adjacency(faces)
    keep = np.zeros(n, dtype=bool)
    if len(adj) == 0:
        return keep, np.arange(n)
    g = coo_matrix((np.ones(len(adj)), (adj[:, 0], adj[:, 1])), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    nodes = np.zeros(n, dtype=bool)
    nodes[adj.reshape(-1)] = True
    size = np.bincount(lab, minlength=lab.max() + 1)
    keep = nodes & (size[lab] >= min_len)
    return keep, lab


def clean_mesh(vertices, faces, masks, intrs, c2ws, dilation_radius=11, min_nb_visible=1, upscale=2, min_len=500):
    """clean_mesh (:110-129) -> (vertices, faces) numpy."""
    masks = torch.as_tensor(masks).float()
    if masks.dim() > 3:
        masks = masks.mean(dim=-1)
    V = np.asarray(vertices)
    Fc = np.asarray(faces)
    dil = torch.from_numpy(dilate((masks > 0.5).numpy(), dilation_radius))
    valid = vertex_valid(V, dil, intrs, c2ws, min_nb_visible).numpy()
    Fc = Fc[valid[Fc].all(axis=-1)]
    Vf = torch.as_tensor(V).float().numpy()
    nv, h, w = masks.shape
    all_idx = []
    for i in range(nv):
        ro, rd = camera_rays(intrs[i], c2ws[i], h, w, upscale)
        m = F.interpolate(masks[i][None, None], scale_factor=upscale, mode="nearest")[0, 0]
        sel = (m > 0).view(-1).numpy()
        idx, _ = first_hits(Vf, Fc, ro.numpy()[sel], rd.numpy()[sel])
        all_idx.append(np.unique(idx))
    values = sorted(set(np.concatenate(all_idx).tolist()))
    hull = np.zeros(len(Fc), dtype=bool)
    hull[values[1:]] = True
    Fc = Fc[hull]
    keep, _ = components_keep(Fc, min_len)
    Fc = Fc[keep]
    used = np.zeros(len(V), dtype=bool)
    used[Fc.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    return V[used], remap[Fc]
