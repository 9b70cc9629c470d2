"""Row-wise sharding of a corpus across ranks: contiguous, screening-tile aligned
blocks; global row id = base + local row.  Pure host logic shared by bench.py and the tests."""

TILE_ROWS = 256


def shard_range(rows, world, rank):
    """-> (base, n_local) of `rank`'s block; blocks are disjoint and cover [0, rows)."""
    per = (rows + world - 1) // world
    per = (per + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS
    base = min(rank * per, rows)
    return base, max(0, min(rows, base + per) - base)
