"""world_size-2 (and 3, ragged) gloo test of the multi-rank host logic on CPU: every rank
renders its row tile (contiguous, and interleaved as the fused peer-store path does), the rows are
gathered, and every rank ends with the frame a single process renders.  The tile renderer here is the oracle (there is no GPU in this container);
the -m gpu suite repeats the equality with the CUDA kernel (test_gpu_parity.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, W, H, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from curvis_b200 import scenes
        from curvis_b200.distributed import all_gather_frame, row_tile
        from oracle import oracle as O
        cam = O.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
        bp, bn = scenes.noise_background(256, 128, 1), scenes.noise_background(256, 128, 2)
        g, s = O.metric("ellis"), O.sim(200, 10.0, 0.1)
        b, e = row_tile(H, rank, world)
        tile, _, st = O.render_rows(g, cam, s, bp, bn, row_begin=b, row_end=e, with_records=False)
        frame = torch.empty(H * W * 3, dtype=torch.uint8)
        all_gather_frame(torch.from_numpy(tile.reshape(-1)), frame, H, W)
        steps = torch.tensor([st["total_steps"]], dtype=torch.int64)
        dist.all_reduce(steps)
        full, _, fst = O.render_rows(g, cam, s, bp, bn, with_records=False)
        ok = bool((frame.numpy().reshape(H, W, 3) == full).all()) and int(steps.item()) == fst["total_steps"]
        # interleaved ownership (what the fused peer-store path of bench.py uses): rows rank, rank+world, ...
        from curvis_b200.distributed import all_gather_interleaved, interleaved_rows
        b, e, stride = interleaved_rows(H, rank, world)
        mine, _, st2 = O.render_rows(g, cam, s, bp, bn, row_begin=b, row_end=e, row_stride=stride, with_records=False)
        frame2 = torch.zeros(H * W * 3, dtype=torch.uint8)
        all_gather_interleaved(torch.from_numpy(mine.reshape(-1)), frame2, H, W)
        steps2 = torch.tensor([st2["total_steps"]], dtype=torch.int64)
        dist.all_reduce(steps2)
        ok = ok and bool((frame2.numpy().reshape(H, W, 3) == full).all()) and int(steps2.item()) == fst["total_steps"]
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H", [(2, 32, 18), (3, 16, 10)])
def test_row_tiles_all_gather_equals_single_process(built, world, W, H):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_row_tile_partition():
    from curvis_b200.distributed import row_tile
    for H in (1, 7, 144, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            tiles = [row_tile(H, r, world) for r in range(world)]
            assert tiles[0][0] == 0 and tiles[-1][1] == H
            assert all(tiles[i][1] == tiles[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        row_tile(10, 2, 2)


def test_interleaved_rows_partition():
    from curvis_b200.distributed import frame_offset, interleaved_rows
    for H in (1, 7, 144, 2160):
        for world in (1, 2, 3, 8):
            owned = [list(range(*interleaved_rows(H, r, world))) for r in range(world)]
            assert sorted(sum(owned, [])) == list(range(H))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    # more ranks than rows: the surplus ranks own an empty tile whose bounds the library accepts (row_begin <= row_end <= H)
    for H, world in ((1, 8), (3, 8), (5, 7)):
        owned = [interleaved_rows(H, r, world) for r in range(world)]
        assert all(b <= e <= H for b, e, _ in owned)
        assert sorted(sum((list(range(*o)) for o in owned), [])) == list(range(H))
    assert frame_offset(0, 0, 2160, 3840) == 0 and frame_offset(2, 5, 2160, 3840) == (2 * 2160 + 5) * 3840 * 3
    with pytest.raises(ValueError):
        interleaved_rows(10, 3, 3)


def test_interleaved_blocks_partition():
    """curvis_render_frames_peers_blocks' ownership: every block of every row belongs to exactly one rank, every rank owns an
    equal share (+-1 block) of the frame AND of every single row that has at least `world` blocks."""
    from curvis_b200.distributed import block_owner, interleaved_blocks
    for H, W, bw in ((9, 160, 32), (2160, 3840, 64), (4320, 7680, 64), (7, 96, 96), (5, 100, 64)):
        for world in (1, 2, 3, 8):
            tiles = [interleaved_blocks(H, W, r, world, bw) for r in range(world)]
            width_b = tiles[0][3]
            assert all(t[3] == width_b for t in tiles) and W % width_b == 0       # (a width that does not divide: whole rows)
            per_row = W // width_b
            owned = [list(range(b, e, s)) for b, e, s, _ in tiles]
            assert sorted(sum(owned, [])) == list(range(H * per_row))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
            if per_row >= world and H <= 16:
                for row in range(H):
                    counts = [sum(1 for v in o if v // per_row == row) for o in owned]
                    assert max(counts) - min(counts) <= 1, (H, W, bw, world, row)
            for r, o in enumerate(owned[:2]):
                for v in o[:5]:
                    assert block_owner(v // per_row, (v % per_row) * width_b, W, world, width_b) == r
    assert interleaved_blocks(3, 64, 7, 8, 64) == (3, 3, 8, 64)                  # surplus rank: an empty tile
    with pytest.raises(ValueError):
        interleaved_blocks(10, 64, 3, 3)
