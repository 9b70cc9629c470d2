"""CPU tests of host-side logic that needs no GPU."""
import numpy as np
import pytest


def test_merge_reverse_keeps_own_picks_then_reverse_edges():
    from surrealdb_b200.hnsw_build import _merge_reverse
    # 4 nodes, m_max = 2; forward picks (nearest first)
    fwd = np.array([[1, 2], [0, -1], [0, 3], [0, 1]], np.int64)
    cnt = np.array([2, 1, 2, 2], np.int64)
    out, n = _merge_reverse(fwd, cnt, m_max=2, cap=6)
    rows = [list(out[i, : n[i]]) for i in range(4)]
    assert rows[0][:2] == [1, 2] and set(rows[0][2:]) == {3}            # 0 is picked by 1, 2, 3; 1 and 2 are duplicates
    assert rows[1] == [0, 3]                                            # own pick, then the reverse edge from 3
    assert rows[2] == [0, 3]                                            # 0 -> 2 duplicates the forward edge 2 -> 0
    assert rows[3][:2] == [0, 1] and set(rows[3][2:]) == {2}
    # reverse edges beyond the capacity are dropped, own picks never are
    fwd = np.zeros((40, 1), np.int64)
    fwd[0, 0] = 1
    cnt = np.ones(40, np.int64)
    out, n = _merge_reverse(fwd, cnt, m_max=4, cap=8)
    assert 4 <= n[0] <= 5 and out[0, 0] == 1                            # 39 nodes point at 0: own pick + <= 4 reverse slots
    assert all(out[i, 0] == 0 for i in range(2, 40))


def test_level_law_matches_the_reference_distribution():
    from surrealdb_b200.hnsw_build import assign_levels
    lv = assign_levels(200_000, 16, seed=3)
    # P(level >= l) = m^-l  (hnsw/mod.rs:263-266)
    for l in (1, 2, 3):
        frac = (lv >= l).mean()
        assert abs(frac - 16.0 ** -l) < 4 * np.sqrt(16.0 ** -l / lv.size) + 1e-4


def test_varint_and_state_reject_garbage():
    from surrealdb_b200 import staging as S
    assert S.read_varint(b"\x05", 0) == (5, 1)
    assert S.read_varint(b"\xfb\x00\x03", 0) == (768, 3)
    assert S.read_varint(b"\xfc\x00\x00\x01\x00", 0) == (65536, 5)
    with pytest.raises(ValueError):
        S.read_varint(b"\xfb\x00", 0)
    with pytest.raises(ValueError):
        S.read_varint(b"\xfe" + b"\0" * 16, 0)
    with pytest.raises(ValueError):
        S.parse_hnsw_state(b"\x02\x00\x00\x01\x00\x00\x00")


def test_pack_values_layout():
    from surrealdb_b200.staging import pack_values
    blob, off, ids = pack_values([(7, b"abc"), (9, b""), (2, b"de")])
    assert bytes(blob) == b"abcde" and off.tolist() == [0, 3, 3, 5] and ids.tolist() == [7, 9, 2]
    blob, off, ids = pack_values([])
    assert off.tolist() == [0] and ids.size == 0
