"""The reference's tile arithmetic (Sampling.hs:5-21, main.hs:109-131) as used for sharding.

The data-parallel unit of the reference is one `squareBatches` window per spark / Par task
(main.hs:82,93).  The multi-GPU driver assigns those same windows to ranks, strided:
tile i -> rank i mod G.  The arithmetic itself lives in the C library
(yahr_b200_num_batches / yahr_b200_batch_window); this module only enumerates.
"""
from . import api


def tile_windows(width, height, num_threads=1):
    """All windows (x0, y0, x1, y1) of squareBatches width height nBatches, in batch order."""
    n = api.num_batches(num_threads, width, height)
    return [api.batch_window(width, height, i, n) for i in range(n)]


def rank_tiles(width, height, world_size, rank, num_threads=1):
    """Indices of the tiles rank `rank` of `world_size` renders."""
    n = api.num_batches(num_threads, width, height)
    return list(range(rank, n, world_size))


def rank_pixel_count(width, height, world_size, rank):
    wins = tile_windows(width, height)
    return sum((x1 - x0) * (y1 - y0) for (x0, y0, x1, y1) in wins[rank::world_size])


def tile_grid(width, height, num_threads=1):
    """(nx, ny) of squareBatches' tile grid: (nx, ny) = loop count 1 (Sampling.hs:11-15); tile `num` sits in row
    num // nx (Sampling.hs:16)."""
    n = api.num_batches(num_threads, width, height)
    nx, ny = n, 1
    while nx % 2 == 0 and 2 * width * ny < height * nx:
        nx //= 2
        ny *= 2
    return nx, ny


def shard_block_rows(width, height, world_size):
    """Rows per block of the host-buffer shards (yahr_b200_render_shard; capi.cu tilesFor, byRows = 2): up to four
    consecutive tile rows, at least eight blocks per shard."""
    _, ny = tile_grid(width, height)
    return max(1, min(4, ny // (8 * world_size)))


def rank_rows(width, height, world_size, rank, block=1):
    """Row sharding: the tile rows rank `rank` renders and the pixel rows [y0, y1) each covers.  The rows are dealt in
    blocks of `block` consecutive rows, block b -> rank b mod G: block = 1 for the "rows" exchange
    (yahr_b200_render_device_shard), shard_block_rows() for the host-buffer shards (yahr_b200_render_shard)."""
    nx, ny = tile_grid(width, height)
    n = nx * ny
    out = []
    for r in range(ny):
        if (r // block) % world_size != rank:
            continue
        ys = [api.batch_window(width, height, r * nx + i, n) for i in range(nx)]
        ys = [(y0, y1) for (x0, y0, x1, y1) in ys if x1 > x0 and y1 > y0]
        if ys:
            out.append((r, ys[0][0], ys[0][1]))
    return out
