"""Multi-GPU sharding of Renderer::sample.

The reference parallelises over image rows with rayon (src/renderer.rs:118-127); pixels
and samples are independent and the scene is read-only.  Here the image is cut into
16x8-pixel tiles dealt round-robin to ranks (tile t belongs to rank t % world), every
rank (one process per GPU) renders only its tiles into a zero-initialised full-size
float3 buffer, and ONE all-reduce(sum) over NCCL assembles the image.  Because the RNG
stream is keyed by (seed, pixel, sample) and every pixel is summed by exactly one rank,
the result is bit-identical for any world size (x + 0 is exact).

Two ways to assemble, same bits:

  * `assemble` / `render_distributed`: every rank renders into a zeroed FULL-size buffer, one all-reduce(sum).  N x the
    image in flight.
  * `gather_tiles` / `render_distributed_gather`: every rank renders ONLY its own tiles into a compact tile-major
    buffer (rptb_render_params.compact_out, 1/N of the image), one all-gather, then a fixed permutation puts the
    pixels in row-major order.  1 x the image in flight, and nothing is summed at all -- what bench.py times.

There is no other exchange step on this path, so no other collective.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _capi as capi

TILE_W, TILE_H = 16, 8  # must match rpt_b200/csrc/integrator.cuh


def tile_owner(width: int, height: int, shard_count: int) -> np.ndarray:
    """(height, width) int array: which shard renders each pixel."""
    tiles_x = (width + TILE_W - 1) // TILE_W
    ys, xs = np.mgrid[0:height, 0:width]
    tile = (ys // TILE_H) * tiles_x + (xs // TILE_W)
    return (tile % shard_count).astype(np.int32)


def shard_tiles(width: int, height: int, shard_index: int, shard_count: int) -> int:
    """How many 16x8 tiles shard `shard_index` of `shard_count` owns (RenderArgs::ntiles_mine)."""
    ntiles = ((width + TILE_W - 1) // TILE_W) * ((height + TILE_H - 1) // TILE_H)
    return (ntiles - shard_index + shard_count - 1) // shard_count if ntiles > shard_index else 0


def gather_permutation(width: int, height: int, shard_count: int) -> np.ndarray:
    """For the concatenation of the shards' compact buffers, each padded to shard 0's size (the largest):
    perm[y * width + x] = index of that pixel in the concatenation (in pixels, not floats).  Mirrors rptb_tile_pixel."""
    tiles_x = (width + TILE_W - 1) // TILE_W
    per = shard_tiles(width, height, 0, shard_count) * TILE_W * TILE_H
    ys, xs = np.mgrid[0:height, 0:width]
    tile = (ys // TILE_H) * tiles_x + (xs // TILE_W)
    lx, ly = xs % TILE_W, ys % TILE_H
    j = ((ly // 4) * 2 + lx // 8) * 32 + (ly % 4) * 8 + lx % 8   # warp (ly/4, lx/8) covers 8x4 pixels, lane = row-major inside it
    return ((tile % shard_count) * per + (tile // shard_count) * (TILE_W * TILE_H) + j).astype(np.int64).ravel()


def gather_tiles(render_compact: Callable[[int, int], "object"], width: int, height: int, perm=None, group=None):
    """All ranks call this.  `render_compact(rank, world)` -> 1-D float tensor holding this rank's tiles, tile-major
    (shard_tiles(...) * 384 values; it may be longer -- it is cut / padded to shard 0's size).  One all-gather, then
    the pixels are put in row-major order: returns (height*width, 3) on every rank."""
    import torch
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    mine = render_compact(rank, world)
    per = shard_tiles(width, height, 0, world) * TILE_W * TILE_H * 3
    if mine.numel() != per:
        padded = mine.new_zeros(per)
        padded[:min(per, mine.numel())] = mine[:per]
        mine = padded
    if world > 1:
        allv = mine.new_empty(per * world)
        dist.all_gather_into_tensor(allv, mine, group=group)
    else:
        allv = mine
    if perm is None:
        perm = torch.from_numpy(gather_permutation(width, height, world)).to(allv.device)
    return allv.view(-1, 3).index_select(0, perm)


def assemble(render_shard: Callable[[int, int], "object"], group=None):
    """The collective step, independent of what renders a shard: every rank calls
    `render_shard(rank, world)` -> a tensor holding its tiles and zeros elsewhere, then one
    all-reduce(sum).  Used with the CUDA shard renderer in production (NCCL) and with a CPU
    shard renderer in the gloo tests."""
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    out = render_shard(rank, world)
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def render_shard_device(renderer, iterations: int, out, shard_index: int, shard_count: int, first_sample: int = 0,
                        stream: Optional[int] = None, stats: Optional[capi.Stats] = None, collect_stats: int = 0,
                        compact: bool = False) -> None:
    """Launch this shard's part of Renderer::sample into `out`, a CUDA float32 tensor of
    width*height*3 elements (torch) on the renderer's device.  `stream` is a raw
    cudaStream_t; torch's default stream has handle 0, which is passed as cudaStreamLegacy
    (0x1) because NULL means "the library's own stream, synchronous" at the C ABI.  compact = True: `out` holds only
    this shard's tiles, tile-major (shard_tiles(...) * 384 floats; rptb_render_params.compact_out)."""
    ds = renderer.device_scene()
    p = renderer.params(iterations, first_sample, shard_index, shard_count, collect_stats)
    p.compact_out = 1 if compact else 0
    cam = renderer.camera.to_c()
    capi.check(
        capi.lib().rptb_render_samples_device(ds.handle, C.byref(cam), C.byref(p), C.c_void_p(out.data_ptr()),
                                              C.c_void_p(stream or 1), C.byref(stats) if stats is not None else None),
        "rptb_render_samples_device",
    )


def render_distributed(renderer, iterations: int, first_sample: int = 0, group=None, out=None):
    """All ranks call this; returns the full image as a CUDA float32 tensor (H*W, 3) on
    every rank.  One process per GPU (torchrun); NCCL over NVLink.  The kernel and the
    all-reduce are enqueued on the same stream: no host synchronisation in between."""
    import torch

    dev = torch.device("cuda", renderer._first_device())
    if out is None:
        out = torch.empty(renderer._width * renderer._height * 3, dtype=torch.float32, device=dev)

    def shard(rank: int, world: int):
        stream = torch.cuda.current_stream(dev).cuda_stream
        render_shard_device(renderer, iterations, out, rank, world, first_sample, stream)
        return out

    return assemble(shard, group).view(-1, 3)


def render_distributed_gather(renderer, iterations: int, first_sample: int = 0, group=None, scratch=None, perm=None):
    """Like render_distributed, through the all-gather of compact shards: (H*W, 3) float32 on every rank."""
    import torch

    dev = torch.device("cuda", renderer._first_device())
    w, h = renderer._width, renderer._height

    def shard(rank: int, world: int):
        n = shard_tiles(w, h, 0, world) * TILE_W * TILE_H * 3
        buf = scratch if scratch is not None and scratch.numel() == n else torch.zeros(n, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        render_shard_device(renderer, iterations, buf, rank, world, first_sample, stream, compact=True)
        return buf

    return gather_tiles(shard, w, h, perm, group)
