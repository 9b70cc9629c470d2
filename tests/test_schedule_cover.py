"""The screen's tile schedules visit EVERY corpus tile exactly once (host logic of api.cu, no GPU needed).

A row whose tile is never visited can never become a candidate, and the exactness proof assumes every row was screened,
so a hole in a schedule would be a silent wrong answer.  The schedules are plain host code behind the diagnostic entry
point sdb_debug_schedule: the streaming schedule (probe tiles + ONE main launch in a golden-ratio permuted order) and the
multi-pass schedule (interleaved strided subsets).  Checked here for every tile count up to a few thousand and for a
spread of larger ones, for both candidate capacities and both probe sizes."""
import ctypes as C

import numpy as np
import pytest

from surrealdb_b200 import _lib as L

TILE = 256


def schedule(n_rows, cap, k, nq, streaming):
    lib = L.lib()
    n, npb = C.c_uint64(), C.c_uint32()
    L.check(lib.sdb_debug_schedule(C.c_uint64(n_rows), C.c_uint32(cap), C.c_uint32(k), C.c_uint32(nq), C.c_int(streaming),
                                   None, C.c_uint64(0), C.byref(n), None, C.c_uint32(0), C.byref(npb)))
    tiles = np.zeros(max(int(n.value), 1), np.uint32)
    probe = np.zeros(max(int(npb.value), 1), np.uint32)
    L.check(lib.sdb_debug_schedule(C.c_uint64(n_rows), C.c_uint32(cap), C.c_uint32(k), C.c_uint32(nq), C.c_int(streaming),
                                   tiles.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(tiles.size), C.byref(n),
                                   probe.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(probe.size), C.byref(npb)))
    return tiles[: n.value], probe[: npb.value]


def check_cover(n_rows, cap, k, nq, streaming):
    T = (n_rows + TILE - 1) // TILE
    tiles, probe = schedule(n_rows, cap, k, nq, streaming)
    assert tiles.size == T, (n_rows, cap, k, streaming, tiles.size, T)
    if T:
        seen = np.bincount(tiles, minlength=T)
        assert seen.size == T and (seen == 1).all(), (n_rows, cap, k, streaming)
        if probe.size:
            assert probe.max() < T and np.unique(probe).size == probe.size, (n_rows, cap, k)
            assert probe.size <= 64 and probe.size * 8 <= 512  # the probe buffer holds 512 chunk maxima per query


@pytest.mark.parametrize("streaming", [1, 0])
def test_every_tile_count_up_to_3000_is_covered_exactly_once(streaming):
    for T in range(0, 3000):
        n_rows = T * TILE - (T % 3) * 17 if T else 0  # ragged last tiles too
        n_rows = max(n_rows, 0)
        for cap, k, nq in ((4096, 10, 1024), (16384, 100, 256)):
            check_cover(n_rows, cap, k, nq, streaming)


@pytest.mark.parametrize("streaming", [1, 0])
def test_large_corpora_are_covered_exactly_once(streaming):
    rng = np.random.default_rng(5)
    sizes = [10_000_000, 1_250_000, 2_500_000, 5_000_000, 39063 * TILE, 4883 * TILE, 100_000_000] + \
            [int(x) for x in rng.integers(3000 * TILE, 60_000_000, 40)]
    for n_rows in sizes:
        for cap, k, nq in ((4096, 10, 1024), (4096, 100, 4096), (16384, 10, 8)):
            check_cover(n_rows, cap, k, nq, streaming)


def test_streaming_order_is_a_spread_permutation():
    # the main launch must sample the whole corpus early (sorted / clustered corpora): among the first 1 % of the visits
    # of a 10M-row corpus every tenth of the corpus appears
    tiles, probe = schedule(10_000_000, 4096, 10, 1024, 1)
    T = tiles.size
    first = tiles[: T // 100]
    assert np.unique(first // (T // 10 + 1)).size == 10
    assert probe.size == 16 and np.unique(probe // (T // 16)).size == 16
