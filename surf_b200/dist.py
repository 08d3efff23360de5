"""Ray / grid sharding across ranks and the all-gather of results.

Rays and grid points are independent given the read-only scene, so the path has no exchange step:
ranks render disjoint ray ranges (borders on 256-ray chunk boundaries, so the per-chunk RNG stream of
quirk Q1 and the empty-mask fallback of Q6 are identical to the single-GPU run) or disjoint x-slabs of
the SDF grid, and one all-gather assembles the image / grid.  Works with any torch.distributed backend
(NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import queue
import threading
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


def render_image_sharded(module, scene, rays_o, rays_d, near, far, cos_anneal_ratio=1.0, chunk=256, t_rand=None,
                         jitter=None):
    """validate()'s image pass split over the ranks of the default process group + all-gather.
    Every rank draws the same host jitter stream (same seed) and uses its slice; `jitter`: a JitterPrefetcher of this
    rank's shard that draws it a step ahead on a worker thread instead."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = rays_o.shape[0]
    r0, r1 = shard_rays(n, rank, world, chunk)
    if jitter is not None:
        mine = jitter.next()
    else:
        if t_rand is None and module.perturb > 0:
            mine = draw_shard_randoms(module, n, r0, r1, chunk)
        else:
            mine = None if t_rand is None else t_rand[r0:r1]
    if near.shape[0] != 1:
        near, far = near[r0:r1], far[r0:r1]
    res = module.render_image(scene, rays_o[r0:r1], rays_d[r0:r1], near, far, cos_anneal_ratio, chunk, mine)
    return gather_image(res, n, chunk)


def skip_cpu_rng(n_draws: int) -> None:
    """Advance torch's global CPU generator by ``n_draws`` float32 draws without producing them (== discarding
    ``torch.rand(n_draws)``, bit for bit, ~7x faster: only the mt19937 state transition runs, in the C library)."""
    import ctypes as C

    import numpy as np

    from . import _lib
    if n_draws <= 0:
        return
    st = torch.get_rng_state()
    if st.numel() != 5056:
        torch.rand(int(n_draws))            # an unknown state layout: draw and discard
        return
    buf = st.numpy()                        # CPUGeneratorImplStateLegacy: seed u64 | left i32 | seeded i32 | next u64 | state[624] u64
    if int(buf[12:16].view(np.int32)[0]) == 0:
        torch.rand(1)                       # not seeded yet: let torch seed it, then skip the rest
        return skip_cpu_rng(n_draws - 1)
    base = buf.ctypes.data
    _lib.check(_lib.load().surf_mt19937_skip(C.c_void_p(base + 24), C.c_void_p(base + 8), C.c_void_p(base + 16),
                                             C.c_uint64(int(n_draws))), "mt19937_skip")
    torch.set_rng_state(st)


def _chunk_draws(module, n_rays: int, chunk: int) -> int:
    """Draws the chunks of ``n_rays`` rays consume (ImplicitSurface.draw_chunk_randoms)."""
    from .modules.implicit_surface import N_RANDOM_PTS
    n_st = len(module.n_samples)
    full, rem = divmod(n_rays, chunk)
    return full * (n_st * chunk + N_RANDOM_PTS * 3) + ((n_st * rem + N_RANDOM_PTS * 3) if rem else 0)


def draw_shard_randoms(module, n_rays: int, r0: int, r1: int, chunk: int = 256) -> torch.Tensor:
    """Jitter table (r1-r0, n_stages) of the rays [r0, r1) of an n_rays image, leaving the global CPU generator where
    the single-process image pass leaves it: the draws of the other chunks are skipped, not produced."""
    assert (r0 % chunk == 0 or r0 == n_rays) and (r1 % chunk == 0 or r1 == n_rays), "shard borders must sit on chunk boundaries"
    skip_cpu_rng(_chunk_draws(module, r0, chunk))
    t = module.draw_chunk_randoms(r1 - r0, chunk) if r1 > r0 else torch.empty((0, len(module.n_samples)))
    skip_cpu_rng(_chunk_draws(module, n_rays - r1, chunk))
    return t


class JitterPrefetcher:
    """The jitter tables of consecutive images for ONE rank's ray shard, drawn a step ahead on a worker thread.

    The reference draws its jitter from torch's global CPU generator, sequentially over the image (quirk Q1): 4 x
    rand([256,1]) + rand([1024,3]) per 256-ray chunk, 7.4 M draws per 576x800 image.  To render a shard
    bit-identically to the single-GPU image a rank has to consume the draws of the chunks before its shard, draw its
    own, and consume the rest so that the NEXT image starts at the right place of the stream.  Producing all of them
    with torch.rand is ~20 ms of host time per image on every rank — more than the 15 ms an 8-GPU image takes on the
    device; draw_shard_randoms skips the foreign chunks in the C library (state transition only, ~3 ms) and a worker
    thread does even that for image k+1 while the GPU renders image k; `next()` hands out a pinned (r1-r0, n_stages)
    table.  The worker is the only consumer of the global generator while the prefetcher is open (seed before creating
    it; close() it before drawing anything else)."""

    def __init__(self, module, n_rays: int, r0: int, r1: int, chunk: int = 256, depth: int = 2):
        assert (r0 % chunk == 0 or r0 == n_rays) and (r1 % chunk == 0 or r1 == n_rays), "shard borders must sit on chunk boundaries"
        self._m, self._n, self._r0, self._r1, self._chunk = module, int(n_rays), int(r0), int(r1), int(chunk)
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._stop = threading.Event()
        self._err = None
        self._pin = torch.cuda.is_available()
        self._t = threading.Thread(target=self._work, name="surf-jitter", daemon=True)
        self._t.start()

    def _work(self):
        try:
            while not self._stop.is_set():
                t = draw_shard_randoms(self._m, self._n, self._r0, self._r1, self._chunk)
                if self._pin:
                    t = t.pin_memory()
                while not self._stop.is_set():
                    try:
                        self._q.put(t, timeout=0.05)
                        break
                    except queue.Full:
                        continue
        except BaseException as e:      # surfaced by next()
            self._err = e

    def next(self) -> torch.Tensor:
        while True:
            if self._err is not None:
                raise RuntimeError("jitter worker failed") from self._err
            try:
                return self._q.get(timeout=0.05)
            except queue.Empty:
                continue

    def close(self):
        self._stop.set()
        self._t.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
