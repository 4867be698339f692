"""world_size-2 (and 3) gloo tests of the N>1 host logic on CPU: strided tile assignment and the
end-of-frame exchange.  The per-rank pixels come from the CPU oracle rendering exactly the tile
subset the rank would render on its GPU (tile i -> rank i mod G); the exchange is the same
`reduce(sum)` of zero-initialised frames that TileShardedRenderer's "reduce" mode issues over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as ob
from yahr_b200 import api, scenes, tiles
from yahr_b200.dist import gather_tiles_reference


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, cam = scenes.c1_scene_yahrr(200, 150)
    w, h = api.image_size(cam)
    o = ob.OracleScene(sc)
    rgb = np.zeros((h, w, 3), np.float32)
    pid = np.zeros((h, w), np.uint32)
    tt = np.zeros((h, w), np.float32)
    o.render(cam, tile_stride=world, tile_offset=rank, threads=1, out=(rgb, pid, tt))
    # this rank wrote exactly the pixels of its tiles
    mine = np.zeros((h, w), bool)
    wins = tiles.tile_windows(w, h)
    for i in tiles.rank_tiles(w, h, world, rank):
        x0, y0, x1, y1 = wins[i]
        mine[y0:y1, x0:x1] = True
    assert (rgb[~mine] == 0).all()
    frame = torch.from_numpy(rgb.copy())
    ids = torch.from_numpy(pid.view(np.int32).copy())
    dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(ids, dst=0, op=dist.ReduceOp.SUM)
    # also exercise the all-gather route used for cross-checking
    parts = [torch.zeros_like(frame) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(rgb))
    if rank == 0:
        full_rgb, full_pid, _, _ = o.render(cam, threads=1)
        ok1 = np.array_equal(frame.numpy().view(np.uint32), full_rgb.view(np.uint32))
        ok2 = np.array_equal(ids.numpy().view(np.uint32), full_pid)
        ok3 = np.array_equal(gather_tiles_reference(w, h, world, [p.numpy() for p in parts]).view(np.uint32),
                             full_rgb.view(np.uint32))
        q.put((ok1, ok2, ok3))
    o.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_sharding_and_reduce_exchange_gloo(world):
    api.build_library()
    ob.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == (True, True, True)


def test_rank_tiles_partition():
    api.build_library()
    for (w, h) in [(512, 512), (1920, 1080), (37, 11)]:
        n = api.num_batches(1, w, h)
        for world in (1, 2, 4, 8):
            seen = sorted(t for r in range(world) for t in tiles.rank_tiles(w, h, world, r))
            assert seen == list(range(n))
            assert sum(tiles.rank_pixel_count(w, h, world, r) for r in range(world)) == w * h


def test_rank_rows_partition():
    """Row sharding: every pixel row belongs to exactly one rank, rows of a rank are whole tile rows."""
    api.build_library()
    for (w, h) in [(512, 512), (1920, 1080), (3840, 2160), (37, 11), (5, 300)]:
        for world in (1, 2, 3, 8):
            for block in (1, tiles.shard_block_rows(w, h, world), 3):
                owner = np.full(h, -1)
                for r in range(world):
                    for (_, y0, y1) in tiles.rank_rows(w, h, world, r, block=block):
                        assert (owner[y0:y1] == -1).all()
                        owner[y0:y1] = r
                assert (owner >= 0).all()
