"""Multi-GPU decomposition (pbrt-rust_b200/distributed.py) driven on CPU: world_size 2 over gloo.

The render callable injected here is the CPU oracle -- the product default is the CUDA library -- so what is
under test is the host logic: tile ownership covers the frame exactly once (static interleave and dynamic
stealing from the shared-memory work counter, super-tile numbering, guided claim sizes -- or from the c10d store counter)
and the single end-of-render film reduce reproduces the one-process render, frame after frame."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, outdir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("pbrt-rust_b200")
    from oracle import oracle as O
    D = importlib.import_module("pbrt-rust_b200.distributed")
    store = dist.TCPStore("127.0.0.1", port, world, rank == 0)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)
    setup = pkg.scenes.small_mixed_scene()
    integ = setup.make_integrator(spp_=2, res=(80, 48))  # 5 x 3 tiles, gaussian filter (footprints straddle tiles)
    film = torch.zeros((integ.film.width * integ.film.height, 4), dtype=torch.float32)
    view = film.numpy()

    def render_tiles(tile_range, interleave, sample_range, tile_order):
        O.render(setup.flat, integ, nthreads=1, tile_range=tile_range, sample_range=sample_range, rgbw=view, tile_interleave=interleave, tile_order=tile_order)

    queue = D.open_shared_queue(pkg.host, integ, dist=dist, name=f"pbrt_b200_test_{port}", min_chunk_tiles=2) if mode == "shm" else None
    # two frames through the same process group and the same counter: the second must not see the first one's exhausted queue
    for frame in range(2):
        film.zero_()
        jobs = D.render_distributed(render_tiles, film, integ, dist=dist, tile_group=2, dynamic=mode != "static", store=store, chunk_tiles=2, queue=queue)
        np.save(os.path.join(outdir, f"jobs_{frame}_{rank}.npy"), np.array([[j[0], j[1]] for j in jobs], np.int64).reshape(-1, 2))
        if rank == 0:
            np.save(os.path.join(outdir, f"film_{frame}.npy"), view)
        dist.barrier()
    if queue is not None:
        queue.counter.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["static", "shm", "store"], ids=["static-interleave", "dynamic-shared-memory-counter", "dynamic-c10d-store"])
def test_two_ranks_reproduce_single_process_render(pkg, oracle, tmp_path, mode):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    setup = pkg.scenes.small_mixed_scene()
    integ = setup.make_integrator(spp_=2, res=(80, 48))
    want, _ = oracle.render(setup.flat, integ, nthreads=1)
    D = importlib.import_module("pbrt-rust_b200.distributed")
    for frame in range(2):
        got = np.load(tmp_path / f"film_{frame}.npy")
        assert np.allclose(got, want, rtol=2e-5, atol=2e-6), frame
        assert np.allclose(got[:, 3], want[:, 3], rtol=1e-5), frame
        if mode != "static":
            chunks = sorted(tuple(c) for r in range(2) for c in np.load(tmp_path / f"jobs_{frame}_{r}.npy").tolist())
            total = integ.n_tile_positions(D.SUPER_TILE) if mode == "shm" else integ.n_tiles()
            assert chunks[0][0] == 0 and chunks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:])), "chunks must cover the tile numbering exactly once"


def test_super_tile_numbering_covers_every_tile_once(pkg, oracle):
    """tile_order = S renumbers the tiles (S x S blocks, positions past the image edge empty); rendering the numbering in two halves
    gives the image of the reference's row-major numbering."""
    setup = pkg.scenes.small_mixed_scene()
    integ = setup.make_integrator(spp_=1, res=(80, 48))  # 5 x 3 tiles
    want, _ = oracle.render(setup.flat, integ, nthreads=2)
    for S in (2, 8):
        n = integ.n_tile_positions(S)
        assert n == ((5 + S - 1) // S) * ((3 + S - 1) // S) * S * S
        acc = np.zeros_like(want)
        oracle.render(setup.flat, integ, nthreads=2, tile_range=(0, n // 3), rgbw=acc, tile_order=S)
        oracle.render(setup.flat, integ, nthreads=2, tile_range=(n // 3, n), rgbw=acc, tile_order=S)
        assert np.allclose(acc, want, rtol=2e-5, atol=2e-6)
        assert np.array_equal(acc[:, 3] > 0, want[:, 3] > 0)


def test_guided_claims_shrink_and_cover(pkg):
    D = importlib.import_module("pbrt-rust_b200.distributed")
    pos, sizes = 0, []
    while pos < 8704:
        n = D.guided_claim(pos, 8704, 8, 64)
        assert 1 <= n <= 8704 - pos
        sizes.append(n); pos += n
    assert pos == 8704 and sizes[0] == 816 and sizes[7] == 816 and sizes[8] < 816  # round 0: 75 % of the frame in 8 equal claims
    assert all(a >= b for a, b in zip(sizes, sizes[1:-1])) and min(sizes[:-1]) == 64 and len(sizes) < 40


def test_work_counter_is_shared_between_processes(pkg, tmp_path):
    """fetch_add from several processes hands out every value exactly once (the counter lives in POSIX shared memory)."""
    import multiprocessing as mp
    name = f"pbrt_b200_test_mp_{os.getpid()}"
    c = pkg.host.WorkCounter(name, create=True)
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_claimer, args=(name, str(tmp_path / f"c{k}.npy"))) for k in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
    got = np.sort(np.concatenate([np.load(tmp_path / f"c{k}.npy") for k in range(3)]))
    assert c.load() == 3 * 2000 * 3 and np.array_equal(got, np.arange(0, 3 * 2000 * 3, 3))
    assert c.fetch_max(10) == 18000 and c.load() == 18000 and c.fetch_max(20000) == 18000 and c.load() == 20000
    c.close()


def _claimer(name, out):
    sys.path.insert(0, str(ROOT))
    pkg = importlib.import_module("pbrt-rust_b200")
    c = pkg.host.WorkCounter(name, create=False)
    np.save(out, np.array([c.fetch_add(3) for _ in range(2000)], np.int64))
    c.close()


def test_static_ownership_partitions_the_tiles(pkg):
    D = importlib.import_module("pbrt-rust_b200.distributed")
    for n_tiles in (1, 7, 64, 8160):
        for world in (1, 2, 4, 8):
            for group in (1, 8):
                owned = [D.owned_tiles(n_tiles, world, r, group) for r in range(world)]
                flat = sorted(t for o in owned for t in o)
                assert flat == list(range(n_tiles))
                if n_tiles >= world * group * 8:
                    sizes = [len(o) for o in owned]
                    assert max(sizes) - min(sizes) <= group
    assert D.tile_interleave(1, 0) is None
    assert D.tile_interleave(4, 3, 8) == (8, 4, 3)
    with pytest.raises(ValueError):
        D.tile_interleave(2, 2)
