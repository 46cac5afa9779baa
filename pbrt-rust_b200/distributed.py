"""Multi-GPU decomposition of SamplerIntegrator::render (src/core/integrator.rs:274-296,392-396).

The reference parallelises over 16x16 image tiles: worker threads pull tiles from ONE shared queue (rayon `par_iter`,
integrator.rs:291-296) and the film tiles are merged on the main thread (:392-396).  Here the workers are GPUs, one
process each:

  * the scene is replicated on every GPU;
  * `dynamic=True` (the reference's scheme): the queue head is a 64-bit counter in POSIX shared memory
    (`pbrt_b200_work_counter`, host.WorkCounter) -- a claim is one lock-free fetch-add on host memory, no network round
    trip, no server rank.  Tiles are numbered super-tile major (`tile_order` = 8: 128x128-pixel blocks), so a claimed
    range is a compact image region (coherent rays), and claims shrink as the frame runs out (guided self-scheduling:
    a GPU wants tens of millions of samples per call, the tail wants small units), so that all ranks finish together
    whatever the cost distribution over the image;
  * `dynamic=False`: static ownership of interleaved tile groups (`tile_group` consecutive tiles per group, group g
    belongs to rank g % world) -- deterministic, no shared state, balanced when cost varies smoothly over the image;
  * every rank accumulates into a full-frame {r,g,b,w} film (filter footprints may straddle tile ownership), and the
    films are summed onto rank 0 with ONE collective at the end (`torch.distributed.reduce`: NCCL over NVLink on GPUs,
    gloo in the CPU tests) -- the only communication on the data path.

The render callable is injected so the CPU tests can drive this logic with gloo and the oracle; the product default is
the CUDA library through `Scene.render` (there is no CPU fallback).
"""
from __future__ import annotations

import os

SUPER_TILE = 8  # tiles per super-tile edge of the dynamic numbering: 8 x 16 = 128 pixels


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


def guided_claim(position: int, total: int, world: int, min_chunk: int, fraction: float = 0.75):
    """Size of the claim that starts at `position` ("factoring" self-scheduling): the frame is handed out in rounds, round k
    offers `fraction` of what the earlier rounds left, in `world` equal claims -- 75 % of the frame goes out as one large claim per
    rank (a GPU wants tens of millions of samples per call; measured on 2 B200s: eight small claims per rank and step cost 11 %),
    the rest in geometrically shrinking claims that let the ranks finish together.  Never below `min_chunk`, never past the end."""
    import math
    left = max(total - position, 0)
    if left == 0:
        return 0
    done = min(max(position / max(total, 1), 0.0), 1.0 - 1e-12)
    k = int(math.floor(math.log(1.0 - done) / math.log(1.0 - fraction) + 1e-9))
    n = max(int(math.ceil(fraction * total * (1.0 - fraction) ** k / max(world, 1))), int(min_chunk))
    return min(n, left)


class SharedTileQueue:
    """The frame's tile queue: positions [0, total) of a tile numbering, claimed through a shared-memory counter.

    All ranks of the box open the same counter (`name`); rank 0 creates it.  The counter only ever grows: frame f owns
    the positions [f * stride, f * stride + total), stride >= total + slack, so ranks that are still draining frame f can
    never take work of frame f + 1, and nobody has to reset anything between frames (no barrier on the claim path)."""

    def __init__(self, counter, total: int, world: int, min_chunk: int):
        self.counter, self.total, self.world, self.min_chunk = counter, int(total), int(world), int(min_chunk)
        # a late claim overshoots the frame by at most one (first-round) claim per rank
        self.stride = self.total + (self.world + 1) * max(self.total + 1, self.min_chunk)
        self.frame = 0

    def begin_frame(self, frame: int):
        """Entering frame `frame`: whoever arrives first moves the queue head to the frame's base (atomic max).  A rank can only be
        here after it saw frame - 1 exhausted, so no position of an earlier frame is skipped."""
        self.frame = int(frame)
        self.counter.fetch_max(self.frame * self.stride)

    def next_chunk(self):
        base = self.frame * self.stride
        # the claim size depends on the position, which only the fetch-add reveals: read, size, then claim
        seen = self.counter.load()
        pos = max(seen - base, 0)
        if pos >= self.total:
            return None
        n = guided_claim(pos, self.total, self.world, self.min_chunk)
        got = self.counter.fetch_add(n) - base
        if got < 0:
            raise RuntimeError("SharedTileQueue: counter behind the frame base (begin_frame was skipped)")
        if got >= self.total:
            return None
        return got, min(got + n, self.total)


class TileCounter:
    """Dynamic tile stealing through the c10d store of the process group (fallback when no shared-memory counter is given,
    e.g. ranks on different hosts).  Keys are per frame: a second frame in the same group starts from a fresh counter."""

    def __init__(self, store, n_tiles: int, chunk: int, key: str = "pbrt_b200/tiles"):
        self.store, self.n_tiles, self.chunk, self.key = store, int(n_tiles), int(chunk), key

    def next_chunk(self):
        end = self.store.add(self.key, self.chunk)
        begin = end - self.chunk
        if begin >= self.n_tiles:
            return None
        return begin, min(end, self.n_tiles)


_frames = {}  # process-group-wide frame counters for the store fallback (every rank calls render_distributed the same number of times)


def render_distributed(render_tiles, film_tensor, integrator, dist=None, tile_group: int = 8, dynamic: bool = False, store=None,
                       chunk_tiles: int = 256, sample_range=None, job_key: str = "0", queue: SharedTileQueue = None, frame: int = None):
    """Render this rank's share of the frame into `film_tensor` ([npix,4] float32, zeroed by the caller) and sum
    all ranks' films onto rank 0.

    render_tiles(tile_range, tile_interleave, sample_range, tile_order) must ADD filter-weighted samples into film_tensor.
    dynamic + queue: claims from the shared-memory queue (tile numbering `SUPER_TILE`); dynamic + store: c10d store counter
    (row-major numbering); otherwise static interleave.  Returns the list of (begin, end, interleave) jobs this rank rendered.
    """
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n_tiles = integrator.n_tiles()
    jobs = []
    if dynamic and world > 1 and queue is not None:
        if frame is None:
            frame = _frames.get(("shm", job_key), 0)
            _frames[("shm", job_key)] = frame + 1
        queue.begin_frame(frame)
        while True:
            c = queue.next_chunk()
            if c is None:
                break
            render_tiles(c, None, sample_range, SUPER_TILE)
            jobs.append((c[0], c[1], None))
    elif dynamic and world > 1:
        if store is None:
            raise ValueError("dynamic tile stealing needs a SharedTileQueue or the process group's store")
        f = _frames.get(("store", job_key), 0)
        _frames[("store", job_key)] = f + 1
        counter = TileCounter(store, n_tiles, chunk_tiles, key=f"pbrt_b200/tiles/{job_key}/{f}")
        first = True
        while True:
            c = counter.next_chunk()
            if c is None:
                break
            if first and rank == 0 and c[0] != 0 and world == 1:
                raise RuntimeError("tile counter was not fresh")
            first = False
            render_tiles(c, None, sample_range, 0)
            jobs.append((c[0], c[1], None))
    else:
        il = tile_interleave(world, rank, tile_group)
        render_tiles((0, n_tiles), il, sample_range, 0)
        jobs.append((0, n_tiles, il))
    if world > 1:
        dist.reduce(film_tensor, dst=0, op=dist.ReduceOp.SUM)
    return jobs


def open_shared_queue(host, integrator, dist=None, name: str = None, min_chunk_tiles: int = 64):
    """The box-wide tile queue of `integrator`'s frame: rank 0 creates the shared-memory counter, the others attach after a
    barrier.  `host` is the pbrt-rust_b200.host module (it owns the C ABI)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    name = name or f"pbrt_b200_tiles_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}"
    if rank == 0:
        counter = host.WorkCounter(name, create=True)
    if world > 1:
        dist.barrier()
    if rank != 0:
        counter = host.WorkCounter(name, create=False)
    total = integrator.n_tile_positions(SUPER_TILE)
    return SharedTileQueue(counter, total, world, max(min_chunk_tiles, total // (32 * max(world, 1))))
