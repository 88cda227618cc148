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


def interleaved_rows(height: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(row_begin, row_end, row_stride) of rank ``rank`` under interleaved ownership — rows rank, rank+world, ... —
    the arguments of curvis_render_frames_peers.  Every rank gets the same mix of short (sky) and long
    (throat-grazing) rays; a contiguous tile of central rows holds ~3 % more Euler steps than the mean."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    # more ranks than rows: the surplus ranks own an EMPTY tile (row_begin == row_end), like row_tile — a begin beyond the
    # end would be rejected by the library and leave the other ranks waiting at the step barrier
    return min(rank, height), height, world


def interleaved_blocks(height: int, width: int, rank: int, world: int, block_width: int = 64) -> Tuple[int, int, int, int]:
    """(block_begin, block_end, block_stride, block_width) of rank ``rank`` when blocks of ``block_width`` pixels of a row
    are interleaved over the ranks — the arguments of curvis_render_frames_peers_blocks.  Every rank owns a ``world``-th of
    EVERY row, so the two or three rows of a frame that hold its 10^4-step rays are shared by all ranks.  A block width
    that does not divide the frame width falls back to whole rows (block_width = width)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    if block_width <= 0 or width % block_width:
        block_width = width
    n_blocks = height * (width // block_width)
    return min(rank, n_blocks), n_blocks, world, block_width


def block_owner(row: int, col: int, width: int, world: int, block_width: int) -> int:
    """The rank that renders pixel (col, row) under interleaved_blocks."""
    return (row * (width // block_width) + col // block_width) % world


def frame_offset(frame: int, row: int, height: int, width: int) -> int:
    """Byte offset of row ``row`` of frame ``frame`` in a buffer of complete RGB8 frames (frame-major) — where
    curvis_render_frames_peers stores it in every peer's buffer (csrc/geodesic_f64.cuh: finish_ray)."""
    return (frame * height + row) * width * 3


def all_gather_interleaved(rows, frame, height: int, width: int, group=None):
    """The CPU stand-in for the fused peer stores: every rank contributes its interleaved rows (flat uint8
    tensor, its rows in order) and ends with the complete frame, each row at frame_offset(0, row, ...)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    biggest = ((height + world - 1) // world) * width * 3
    padded = torch.zeros(biggest, dtype=rows.dtype, device=rows.device)
    padded[: rows.numel()] = rows
    parts = [torch.empty(biggest, dtype=rows.dtype, device=rows.device) for _ in range(world)]
    if world == 1:
        parts[0].copy_(padded)
    else:
        dist.all_gather(parts, padded, group=group)
    row_bytes = width * 3
    for r in range(world):
        b, e, stride = interleaved_rows(height, r, world)
        for k, row in enumerate(range(b, e, stride)):
            off = frame_offset(0, row, height, width)
            frame[off: off + row_bytes] = parts[r][k * row_bytes: (k + 1) * row_bytes]
    return frame


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
