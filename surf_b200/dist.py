"""Ray / grid sharding across ranks and the all-gather of results.

Rays and grid points are independent given the read-only scene, so the path has no exchange step:
ranks render disjoint ray ranges (borders on 256-ray chunk boundaries, so the per-chunk RNG stream of
quirk Q1 and the empty-mask fallback of Q6 are identical to the single-GPU run) or disjoint x-slabs of
the SDF grid, and one all-gather assembles the image / grid.  Works with any torch.distributed backend
(NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

RECORD_KEYS = ("color_fine", "val_normal", "sdf_depth", "render_depth")   # 3 + 3 + 1 + 1 floats per ray


def shard_rays(n_rays: int, rank: int, world: int, chunk: int = 256) -> Tuple[int, int]:
    """Contiguous [r0, r1) of rank `rank`; every border is a multiple of `chunk`."""
    n_chunks = (n_rays + chunk - 1) // chunk
    per = (n_chunks + world - 1) // world
    r0 = min(n_rays, rank * per * chunk)
    r1 = min(n_rays, (rank + 1) * per * chunk)
    return r0, r1


def shard_planes(resolution: int, rank: int, world: int) -> Tuple[int, int]:
    """x-slab [x0, x1) of the extract_geometry grid (64 planes per rank at 512^3 / 8 GPUs = the
    reference's own block size, implicit_surface.py:338)."""
    per = (resolution + world - 1) // world
    return min(resolution, rank * per), min(resolution, (rank + 1) * per)


def pack_records(res: Dict[str, torch.Tensor]) -> torch.Tensor:
    """(n,8) fp32 record [rgb3, normal3, sdf_depth, render_depth] per ray."""
    return torch.cat([res["color_fine"], res["val_normal"], res["sdf_depth"].reshape(-1, 1),
                      res["render_depth"].reshape(-1, 1)], dim=1).contiguous()


def unpack_records(rec: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"color_fine": rec[:, 0:3], "val_normal": rec[:, 3:6], "sdf_depth": rec[:, 6:7], "render_depth": rec[:, 7]}


def _all_gather_ragged(local: torch.Tensor, sizes: List[int]) -> torch.Tensor:
    """all-gather of row blocks of different length (last shard may be short): pad to the longest."""
    world = dist.get_world_size()
    longest = max(sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def gather_image(local_res: Dict[str, torch.Tensor], n_rays: int, chunk: int = 256) -> Dict[str, torch.Tensor]:
    """All ranks end up with the full-image tensors (14.7 MB at 576x800)."""
    world = dist.get_world_size()
    sizes = [shard_rays(n_rays, r, world, chunk) for r in range(world)]
    sizes = [b - a for a, b in sizes]
    return unpack_records(_all_gather_ragged(pack_records(local_res), sizes))


def gather_grid(local_u: torch.Tensor, resolution: int) -> torch.Tensor:
    """All ranks end up with the full (R,R,R) grid from their x-slabs."""
    world = dist.get_world_size()
    sizes = [shard_planes(resolution, r, world) for r in range(world)]
    sizes = [b - a for a, b in sizes]
    return _all_gather_ragged(local_u.contiguous(), sizes)


def render_image_sharded(module, scene, rays_o, rays_d, near, far, cos_anneal_ratio=1.0, chunk=256, t_rand=None):
    """validate()'s image pass split over the ranks of the default process group + all-gather.
    Every rank draws the same host jitter stream (same seed) and uses its slice."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = rays_o.shape[0]
    if t_rand is None and module.perturb > 0:
        t_rand = module.draw_chunk_randoms(n, chunk)
    r0, r1 = shard_rays(n, rank, world, chunk)
    if near.shape[0] != 1:
        near, far = near[r0:r1], far[r0:r1]
    res = module.render_image(scene, rays_o[r0:r1], rays_d[r0:r1], near, far, cos_anneal_ratio, chunk,
                              None if t_rand is None else t_rand[r0:r1])
    return gather_image(res, n, chunk)
