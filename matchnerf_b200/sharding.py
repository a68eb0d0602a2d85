"""Multi-GPU plumbing for the renderer: one process per GPU, rays partitioned across ranks, ONE all-gather of the
rendered tiles.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport.

Two partitionings are provided:
  * ``row_block``: rank r renders a contiguous block of image rows of ONE target image (single-image latency);
  * image-parallel (``bench.py``): rank r renders target image r of a round of ``world_size`` images.
Either way the exchange step is ``gather_tiles``: every rank contributes its [rays, 5] (rgb, depth, opacity) tile.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def row_block(H: int, W: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_ray, n_rays) of the rows rank ``rank`` owns: rows are split as evenly as possible, earlier ranks
    take the remainder, so blocks are contiguous, disjoint and cover the image in row-major order."""
    base, rem = divmod(H, world)
    r0 = rank * base + min(rank, rem)
    rows = base + (1 if rank < rem else 0)
    return r0 * W, rows * W


def gather_tiles(local: torch.Tensor, counts, group=None) -> torch.Tensor:
    """All-gather ragged [n_r, C] tiles (``counts[r]`` rows from rank r) into the full [sum(counts), C] tensor.
    One collective: tiles are padded to the largest count so ``all_gather_into_tensor`` can be used."""
    world = dist.get_world_size(group)
    pad = max(counts)
    buf = local.new_zeros((pad, local.shape[1]))
    buf[: local.shape[0]] = local
    out = local.new_empty((world * pad, local.shape[1]))
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * pad: r * pad + counts[r]] for r in range(world)], dim=0)


def render_image_sharded(render_range: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                         H: int, W: int, group=None):
    """Render one H x W target image with its rays sharded over the ranks of ``group``.

    ``render_range(first_ray, n_rays)`` -> (rgb [n,3], depth [n] or [n,1], opacity [n] or [n,1]) renders a
    contiguous row-major ray range on the calling rank (e.g. a closure over ``MatchNeRF._render``).
    Returns (rgb [HW,3], depth [HW,1], opacity [HW,1]) on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, n = row_block(H, W, rank, world)
    rgb, depth, opac = render_range(first, n)
    tile = torch.cat([rgb.reshape(n, 3), depth.reshape(n, 1), opac.reshape(n, 1)], dim=1)
    counts = [row_block(H, W, r, world)[1] for r in range(world)]
    full = gather_tiles(tile, counts, group)
    return full[:, :3], full[:, 3:4], full[:, 4:5]
