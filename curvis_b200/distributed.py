"""Row-tile partition of a frame over the ranks of one node and the all-gather that
reassembles it.  Every pixel of ``render_image`` is independent (reference
src/systems.rs:316-326 carries nothing between iterations), so rank g renders rows
[g*H/N, (g+1)*H/N) with no data-path exchange; the only collective is the gather of the RGB8
tiles into the complete frame (torch.distributed: NCCL over NVLink on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def row_tile(height: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [begin, end) of rank ``rank`` — the same split curvis_render_image uses across the
    devices of a context (csrc/curvis_abi.cu)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    return height * rank // world, height * (rank + 1) // world


def all_gather_frame(tile, frame, height: int, width: int, group=None):
    """Gathers every rank's RGB8 row tile (flat uint8 tensor, rows*width*3) into ``frame``
    (flat uint8 tensor, height*width*3) on every rank.  Equal tiles use one
    all_gather_into_tensor; ragged splits (height % world != 0) pad to the largest tile."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        frame.copy_(tile)
        return frame
    row_bytes = width * 3
    if height % world == 0:
        dist.all_gather_into_tensor(frame, tile, group=group)
        return frame
    sizes = [(row_tile(height, r, world)[1] - row_tile(height, r, world)[0]) * row_bytes for r in range(world)]
    biggest = max(sizes)
    padded = torch.zeros(biggest, dtype=tile.dtype, device=tile.device)
    padded[: sizes[rank]] = tile
    parts = [torch.empty(biggest, dtype=tile.dtype, device=tile.device) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    off = 0
    for r in range(world):
        frame[off: off + sizes[r]] = parts[r][: sizes[r]]
        off += sizes[r]
    return frame
