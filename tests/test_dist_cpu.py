"""world_size-2 gloo tests of the shard / all-gather logic (host side; no GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from surf_b200 import dist as sdist


def test_shard_rays_covers_and_aligns():
    for n in (1, 255, 256, 257, 28800, 460800, 460801):
        for world in (1, 2, 3, 4, 8):
            spans = [sdist.shard_rays(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and a <= b
            for a, b in spans:
                assert a % 256 == 0 or a == n
    assert [sdist.shard_planes(512, r, 8) for r in range(8)][3] == (192, 256)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, res, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = {"color_fine": torch.rand(n_rays, 3, generator=g), "val_normal": torch.rand(n_rays, 3, generator=g),
                "sdf_depth": torch.rand(n_rays, 1, generator=g), "render_depth": torch.rand(n_rays, generator=g)}
        r0, r1 = sdist.shard_rays(n_rays, rank, world)
        local = {k: v[r0:r1] for k, v in full.items()}
        got = sdist.gather_image(local, n_rays)
        ok = all(torch.equal(got[k].reshape(full[k].shape), full[k]) for k in full)
        ig = sdist.ImageGather(n_rays, "cpu")
        for _ in range(2):              # persistent buffers: a second image through the same object
            got2 = ig(local)
            ok = ok and all(torch.equal(got2[k].reshape(full[k].shape), full[k]) for k in full)
        u = torch.rand(res, res, res, generator=g)
        x0, x1 = sdist.shard_planes(res, rank, world)
        ok = ok and torch.equal(sdist.gather_grid(u[x0:x1], res), u)
        out_q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_image_and_grid_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    for n_rays, res in ((1000, 9), (1024, 8), (200, 3)):      # ragged, even, and an EMPTY last shard
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, res, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        results = dict(q.get(timeout=10) for _ in range(2))
        assert results == {0: True, 1: True}, (n_rays, res, results)
        port = _free_port()


def test_jitter_prefetcher_reproduces_the_sequential_stream():
    """Every rank's prefetched tables = the rows of the single-process stream over consecutive images (Q1), bit for
    bit, including a ragged last chunk and a rank whose shard is empty."""
    import torch
    from surf_b200 import conf
    from surf_b200.dist import JitterPrefetcher, shard_rays
    from surf_b200.modules.implicit_surface import ImplicitSurface
    m = ImplicitSurface(conf.default_implicit_surface_conf())
    n, images = 256 * 9 + 77, 3
    torch.manual_seed(99)
    want = [m.draw_chunk_randoms(n) for _ in range(images)]
    for world in (1, 3, 16):
        for rank in range(world):
            r0, r1 = shard_rays(n, rank, world)
            torch.manual_seed(99)
            with JitterPrefetcher(m, n, r0, r1) as pf:
                for k in range(images):
                    got = pf.next()
                    assert got.shape == (r1 - r0, 4)
                    assert torch.equal(got, want[k][r0:r1]), (world, rank, k)


def test_skip_cpu_rng_equals_drawing_and_discarding():
    """surf_mt19937_skip advances torch's CPU generator exactly like torch.rand(n) does (state blob and the following
    draws), for counts around the 624-word twist boundaries and for image-sized skips."""
    import time
    import torch
    from surf_b200.dist import skip_cpu_rng
    for seed, pre, n in [(0, 0, 1), (1, 0, 623), (2, 0, 624), (3, 0, 625), (4, 5, 619), (5, 5, 620), (6, 700, 1248),
                         (7, 3, 1_000_003), (8, 0, 7_372_800)]:
        torch.manual_seed(seed)
        if pre:
            torch.rand(pre)
        torch.rand(n)
        want_state, want_next = torch.get_rng_state().clone(), torch.rand(1000)
        torch.manual_seed(seed)
        if pre:
            torch.rand(pre)
        skip_cpu_rng(n)
        assert torch.equal(torch.get_rng_state(), want_state), (seed, pre, n)
        assert torch.equal(torch.rand(1000), want_next), (seed, pre, n)
    torch.manual_seed(0)
    t0 = time.perf_counter(); torch.rand(7_372_800); t1 = time.perf_counter(); skip_cpu_rng(7_372_800); t2 = time.perf_counter()
    print("7.37 M draws: torch.rand %.1f ms, skip %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
