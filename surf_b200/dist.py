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
    """all-gather of row blocks of different length (last shard may be short): pad to the longest.  One collective
    (all_gather_into_tensor) into one buffer; equal shards need no repacking at all."""
    world = dist.get_world_size()
    longest = max(sizes)
    if local.shape[0] == longest:
        pad = local.contiguous()
    else:
        pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    out = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    try:
        dist.all_gather_into_tensor(out, pad)
    except (RuntimeError, NotImplementedError):        # a backend without the flat variant
        bufs = list(out.reshape((world, longest) + tuple(local.shape[1:])).unbind(0))
        dist.all_gather(bufs, pad)
    if all(s_ == longest for s_ in sizes):
        return out
    out = out.reshape((world, longest) + tuple(local.shape[1:]))
    return torch.cat([out[r, :s_] for r, s_ in enumerate(sizes)], dim=0)


def gather_image(local_res: Dict[str, torch.Tensor], n_rays: int, chunk: int = 256) -> Dict[str, torch.Tensor]:
    """All ranks end up with the full-image tensors (14.7 MB at 576x800)."""
    world = dist.get_world_size()
    sizes = [shard_rays(n_rays, r, world, chunk) for r in range(world)]
    sizes = [b - a for a, b in sizes]
    return unpack_records(_all_gather_ragged(pack_records(local_res), sizes))


class ImageGather:
    """gather_image with its buffers allocated once (the per-image call of a render loop): packs the rank's
    (R_local, 8) record block into a persistent padded buffer and all-gathers into a persistent (world, longest, 8)
    buffer — one small pack kernel + ONE NCCL collective per image, no allocation."""

    def __init__(self, n_rays: int, device, chunk: int = 256):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        spans = [shard_rays(n_rays, r, self.world, chunk) for r in range(self.world)]
        self.sizes = [b - a for a, b in spans]
        self.longest = max(self.sizes)
        self.n_rays = n_rays
        self.local = torch.zeros((self.longest, 8), dtype=torch.float32, device=device)
        self.full = torch.empty((self.world * self.longest, 8), dtype=torch.float32, device=device)
        self.even = all(s_ == self.longest for s_ in self.sizes)

    def __call__(self, res: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        n = self.sizes[self.rank]
        loc = self.local
        loc[:n, 0:3] = res["color_fine"]
        loc[:n, 3:6] = res["val_normal"]
        loc[:n, 6:7] = res["sdf_depth"].reshape(-1, 1)
        loc[:n, 7] = res["render_depth"].reshape(-1)
        try:
            dist.all_gather_into_tensor(self.full, loc)
        except (RuntimeError, NotImplementedError):
            dist.all_gather(list(self.full.reshape(self.world, self.longest, 8).unbind(0)), loc)
        if self.even:
            rec = self.full
        else:
            f = self.full.reshape(self.world, self.longest, 8)
            rec = torch.cat([f[r, :s_] for r, s_ in enumerate(self.sizes)], dim=0)
        return unpack_records(rec)


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
