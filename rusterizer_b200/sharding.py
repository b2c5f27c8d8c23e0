"""Multi-GPU partitioning of the raster path (one process per GPU, torch.distributed plumbing).

The path shards in two natural ways (BASELINE.json north_star):
  * independent FRAMES of a camera sweep per GPU -- no data-path collective at all;
  * screen-space TILE-ROW RANGES of one frame per GPU -- geometry is replicated, every rank
    rasterises only its rows (rz_set_row_range) and the resolved strips are gathered (one
    all-gather of equal contiguous strips; NCCL over NVLink on GPUs, gloo in the CPU tests).
Nothing here touches pixels: it is index arithmetic plus one collective call.
"""
from __future__ import annotations


def tile_rows(height: int, tile_h: int) -> int:
    return (height + tile_h - 1) // tile_h


def strip_rows(height: int, world: int, tile_h: int) -> int:
    """Rows per rank: tile rows split evenly (rounded up) so every strip has the same size."""
    return ((tile_rows(height, tile_h) + world - 1) // world) * tile_h


def row_range(rank: int, world: int, height: int, tile_h: int) -> tuple[int, int]:
    """Pixel rows [begin, end) owned by `rank`; tile-aligned except at the bottom edge. May be empty
    (begin == end == height) for trailing ranks of a short image."""
    per = strip_rows(height, world, tile_h)
    return min(height, rank * per), min(height, (rank + 1) * per)


def frame_ids(rank: int, world: int, n_frames: int) -> list[int]:
    """Frames of a sweep rendered by `rank` (round-robin, so neighbouring camera angles spread evenly)."""
    return list(range(rank, n_frames, world))


def gather_strips(strip, height: int, group=None):
    """All-gather equal-sized strips [strip_rows, W] into the full image [height, W] (on every rank).
    `strip` is a torch tensor (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    full = torch.empty((world * strip.shape[0],) + tuple(strip.shape[1:]), dtype=strip.dtype, device=strip.device)
    dist.all_gather_into_tensor(full, strip.contiguous(), group=group)
    return full[:height]
