"""Multi-GPU decomposition of SamplerIntegrator::render (src/core/integrator.rs:274-296,392-396).

The reference parallelises over 16x16 image tiles with one rayon `par_iter` and merges the film tiles on
the main thread.  Here the same tiles are the unit of distribution across GPUs (one process per GPU):

  * the scene is replicated on every GPU;
  * tiles are owned in interleaved groups (`tile_group` consecutive tiles per group, group g belongs to rank
    g % world) -- static, deterministic, and balanced for images whose cost varies smoothly; or handed out
    dynamically in chunks from a shared counter (`TileCounter`, a c10d store `add`) when `dynamic=True`;
  * every rank accumulates into a full-frame {r,g,b,w} film (filter footprints may straddle tile ownership),
    and the films are summed onto rank 0 with ONE collective at the end (`torch.distributed.reduce`, NCCL over
    NVLink on GPUs, gloo in the CPU tests) -- the only communication on the path.

The render callable is injected so the CPU tests can drive this logic with gloo; the product default is the
CUDA library through `Scene.render` (there is no CPU fallback).
"""
from __future__ import annotations


def tile_interleave(world: int, rank: int, tile_group: int = 8):
    """(group, mod, rem) for pbrt_b200_render_desc.tile_group/tile_mod/tile_rem, or None for one rank."""
    if world <= 1:
        return None
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (int(tile_group), int(world), int(rank))


def owned_tiles(n_tiles: int, world: int, rank: int, tile_group: int = 8, tile_begin: int = 0):
    """Tile indices a rank renders under `tile_interleave` (same rule as k_raygen's item -> tile map)."""
    out = []
    for t in range(tile_begin, n_tiles):
        g = (t - tile_begin) // tile_group
        if g % max(world, 1) == (rank if world > 1 else 0):
            out.append(t)
    return out


class TileCounter:
    """Dynamic tile stealing across processes: an atomic counter in the c10d store of the default process group.
    `next_chunk()` returns a [begin, end) range of tiles or None when the frame is exhausted."""

    def __init__(self, store, n_tiles: int, chunk: int, key: str = "pbrt_b200/tiles"):
        self.store, self.n_tiles, self.chunk, self.key = store, int(n_tiles), int(chunk), key

    def next_chunk(self):
        end = self.store.add(self.key, self.chunk)
        begin = end - self.chunk
        if begin >= self.n_tiles:
            return None
        return begin, min(end, self.n_tiles)


def render_distributed(render_tiles, film_tensor, integrator, dist=None, tile_group: int = 8, dynamic: bool = False, store=None,
                       chunk_tiles: int = 256, sample_range=None, job_key: str = "0"):
    """Render this rank's share of the frame into `film_tensor` ([npix,4] float32, zeroed by the caller) and sum
    all ranks' films onto rank 0.

    render_tiles(tile_range, tile_interleave, sample_range) must ADD filter-weighted samples into film_tensor.
    Returns the list of (tile_begin, tile_end, interleave) jobs this rank rendered.
    """
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n_tiles = integrator.n_tiles()
    jobs = []
    if dynamic and world > 1:
        if store is None:
            raise ValueError("dynamic tile stealing needs the process group's store")
        counter = TileCounter(store, n_tiles, chunk_tiles, key=f"pbrt_b200/tiles/{job_key}")
        while True:
            c = counter.next_chunk()
            if c is None:
                break
            render_tiles(c, None, sample_range)
            jobs.append((c[0], c[1], None))
    else:
        il = tile_interleave(world, rank, tile_group)
        render_tiles((0, n_tiles), il, sample_range)
        jobs.append((0, n_tiles, il))
    if world > 1:
        dist.reduce(film_tensor, dst=0, op=dist.ReduceOp.SUM)
    return jobs
