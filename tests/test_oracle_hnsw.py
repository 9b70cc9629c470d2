"""HNSW restatement: structural invariants and the reference's recall tests on its own fixtures
(core/idx/trees/hnsw/mod.rs:1039-1184; fixtures = first rows of tests/data/hnsw-random-*.gz,
committed under tests/golden/ by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _recall(ingest, nq, ext, keep, efs_expect):
    data = np.load(os.path.join(G, "hnsw_ingest_1000x20.f64.npy"))[:ingest].astype(np.float32)
    qs = np.load(os.path.join(G, "hnsw_query_300x20.f64.npy"))[:nq].astype(np.float32)
    # new_params(20, F32, Euclidean, m=8, efc=100, ext, keep, ..)  -> m0 = 2m, ml = 1/ln(m)
    h = O.Hnsw(20, "euclidean", m=8, efc=100, extend_candidates=ext, keep_pruned_connections=keep, seed=42)
    for v in data:
        h.insert(v)
    assert len(h) == ingest and h.check_props()
    for efs, expected in efs_expect:
        total = 0.0
        for q in qs:
            ids, dist = h.search(q, 10, efs)
            b = O.KnnResultBuilder(10)  # add_graph_results: one doc per element (docs = row ids)
            for d, e in zip(dist, ids):
                if b.check_add(d):
                    b.add_graph_result(d, [int(e)])
            res = b.collect()
            assert len(res) == 10
            bi, bd = O.vec_knn_f32(data, q, "euclidean", 10)
            brute = list(zip(bd.tolist(), bi.tolist()))
            rec = len({x[1] for x in res} & {x[1] for x in brute}) / 10.0
            if rec == 1.0:
                assert res == brute  # hnsw/mod.rs:1119-1123
            total += rec
        recall = total / len(qs)
        assert recall >= expected, (efs, recall)


def test_recall_euclidean():
    _recall(1000, 300, False, False, [(10, 0.98), (40, 1.0)])


def test_recall_euclidean_keep_pruned_connections():
    _recall(750, 200, False, True, [(10, 0.98), (40, 1.0)])


def test_recall_euclidean_full():
    _recall(500, 100, True, True, [(10, 0.98), (40, 1.0)])


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "manhattan", "chebyshev"])
@pytest.mark.parametrize("flags", [(False, False), (True, False), (False, True), (True, True)])
def test_small_collections_invariants(metric, flags):
    # tests_hnsw (hnsw/mod.rs:752-791): 30 random vectors, insert then search each -> itself is found
    rng = np.random.default_rng(1)
    data = rng.uniform(-20, 20, (30, 5)).astype(np.float32)
    h = O.Hnsw(5, metric, m=4, efc=500, extend_candidates=flags[0], keep_pruned_connections=flags[1], seed=9)
    for v in data:
        h.insert(v)
        assert h.check_props()
    for i, v in enumerate(data):
        ids, dist = h.search(v, 1, 500)
        assert ids[0] == i and dist[0] <= 1e-6


def test_csr_export_walk_equals_builder_walk():
    rng = np.random.default_rng(2)
    data = rng.uniform(-1, 1, (400, 12)).astype(np.float32)
    h = O.Hnsw(12, "cosine", m=6, efc=60, seed=3)
    for v in data:
        h.insert(v)
    g = h.export()
    assert len(g["layers"]) >= 2
    for q in rng.uniform(-1, 1, (20, 12)).astype(np.float32):
        a = h.search(q, 10, 32)
        vis, exp = h.counters()
        b = O.hnsw_search_csr(g, q, 10, 32)
        assert list(a[0]) == list(b[0]) and list(a[1]) == list(b[1]) and b[2] == (vis, exp)


def test_language_test_hnsw_goldens():
    # language-tests/tests/language/indexes/knn/hnsw_knn_with_condition_new_executor.surql:
    # 7 one-dimensional points, `WHERE flag = true AND point <|2,40|> [44]` -> pts:5 (6), pts:3 (14)
    h = O.Hnsw(1, "euclidean", m=12, efc=150, seed=3)
    for v in (10, 20, 30, 40, 50, 60, 70):
        h.insert(np.array([v], np.float32))
    g = h.export()
    truthy = np.array([1, 0, 1, 0, 1, 0, 1], np.uint8)
    ids, dist, _ = O.hnsw_search_csr(g, np.array([44], np.float32), 2, 40, truthy=truthy)
    assert [int(i) + 1 for i in ids] == [5, 3] and list(dist) == [6.0, 14.0]
    ids, dist, _ = O.hnsw_search_csr(g, np.array([44], np.float32), 2, 40)          # unfiltered: pts:4 (4), pts:5 (6)
    assert [int(i) + 1 for i in ids] == [4, 5] and list(dist) == [4.0, 6.0]
    # reproductions/7229_knn_k_distance_bypasses_hnsw.surql: <|2,100|> [2,3,4,5] -> pts:1 (2), pts:2 (4)
    h = O.Hnsw(4, "euclidean", m=12, efc=500, seed=3)
    for v in ([1, 2, 3, 4], [4, 5, 6, 7], [8, 9, 10, 11]):
        h.insert(np.array(v, np.float32))
    ids, dist = h.search(np.array([2, 3, 4, 5], np.float32), 2, 100)
    assert [int(i) + 1 for i in ids] == [1, 2] and list(dist) == [2.0, 4.0]


def test_simple_hnsw_of_the_reference():
    # idx/trees/hnsw/mod.rs:1001-1037 test_simple_hnsw: 11 two-dimensional points, m = 3 (m0 = 6), efc = 500,
    # extend_candidates + keep_pruned_connections; knn_search((-2,-3), k = 10, ef = 501) must return 10 results
    pts = [(-2, -3), (-2, 1), (-4, 3), (-3, 1), (-1, 1), (-2, 3), (3, 0), (-1, -2), (-2, 2), (-4, -2), (0, 3)]
    for seed in range(8):  # the level assignment is random in the reference too: the property holds for every draw
        h = O.Hnsw(2, "euclidean", m=3, m0=6, efc=500, extend_candidates=True, keep_pruned_connections=True, seed=seed)
        for p in pts:
            h.insert(np.array(p, np.float32))
        assert h.check_props()
        ids, dist = h.search(np.array([-2, -3], np.float32), 10, 501)
        assert len(ids) == 10 and ids[0] == 0 and dist[0] == 0.0
        assert list(dist) == sorted(dist)
        # with ef >= n the walk sees every connected element: the result is the exact top-10
        bi, bd = O.vec_knn_f32(np.array(pts, np.float32), np.array([-2, -3], np.float32), "euclidean", 10)
        assert list(dist) == list(bd)
