"""Multi-GPU decomposition (pbrt-rust_b200/distributed.py) driven on CPU: world_size 2 over gloo.

The render callable injected here is the CPU oracle -- the product default is the CUDA library -- so what is
under test is the host logic: tile ownership covers the frame exactly once (static interleave and dynamic
stealing from the c10d store counter) and the single end-of-render film reduce reproduces the one-process render."""
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


def _worker(rank, world, port, dynamic, outdir):
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
    rendered = []

    def render_tiles(tile_range, interleave, sample_range):
        O.render(setup.flat, integ, nthreads=1, tile_range=tile_range, sample_range=sample_range, rgbw=view, tile_interleave=interleave)
        rendered.append((tile_range, interleave))

    jobs = D.render_distributed(render_tiles, film, integ, dist=dist, tile_group=2, dynamic=dynamic, store=store, chunk_tiles=2)
    np.save(os.path.join(outdir, f"jobs_{rank}.npy"), np.array([[j[0], j[1]] for j in jobs], np.int64))
    if rank == 0:
        np.save(os.path.join(outdir, "film.npy"), view)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dynamic", [False, True], ids=["static-interleave", "dynamic-stealing"])
def test_two_ranks_reproduce_single_process_render(pkg, oracle, tmp_path, dynamic):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, dynamic, str(tmp_path)), nprocs=2, join=True)
    setup = pkg.scenes.small_mixed_scene()
    integ = setup.make_integrator(spp_=2, res=(80, 48))
    want, _ = oracle.render(setup.flat, integ, nthreads=1)
    got = np.load(tmp_path / "film.npy")
    assert np.allclose(got, want, rtol=2e-5, atol=2e-6)
    assert np.allclose(got[:, 3], want[:, 3], rtol=1e-5)
    if dynamic:
        chunks = sorted(tuple(c) for r in range(2) for c in np.load(tmp_path / f"jobs_{r}.npy").tolist())
        nt = integ.n_tiles()
        assert chunks[0][0] == 0 and chunks[-1][1] == nt
        assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:])), "chunks must tile [0, n_tiles) exactly once"


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
