"""KnnTopK selection rule, DoublePriorityQueue and KnnResultBuilder known answers
(core/exec/operators/knn_topk.rs, core/idx/trees/knn.rs:646-788, language-tests knn/*.surql)."""
import numpy as np

from oracle import pyoracle as O


def test_knn_result_builder_kat():
    # idx/trees/knn.rs:647-666 knn_result_builder_test
    b = O.KnnResultBuilder(7)
    b.add_graph_result(0.0, [5])
    b.add_graph_result(0.2, [0, 1, 2])
    b.add_graph_result(0.2, [3])
    b.add_graph_result(0.2, [6, 8])
    assert b.collect() == [(0.0, 5), (0.2, 0), (0.2, 1), (0.2, 2), (0.2, 3), (0.2, 6), (0.2, 8)]


def test_double_priority_queue_kat():
    # idx/trees/knn.rs:734-788 test_double_priority_queue
    q = O.DoublePriorityQueue()
    q.push(2.0, 2); q.push(3.0, 4); q.push(3.0, 3); q.push(1.0, 1)
    assert len(q) == 4 and q.peek_first() == (1.0, 1) and q.peek_last_dist() == 3.0
    assert q.pop_first() == (1.0, 1) and len(q) == 3 and q.peek_first() == (2.0, 2) and q.peek_last_dist() == 3.0
    assert q.pop_first() == (2.0, 2) and len(q) == 2 and q.peek_first() == (3.0, 4) and q.peek_last_dist() == 3.0
    assert q.pop_first() == (3.0, 4) and len(q) == 1 and q.peek_first() == (3.0, 3) and q.peek_last_dist() == 3.0
    assert q.pop_first() == (3.0, 3) and len(q) == 0 and q.peek_first() is None and q.peek_last_dist() is None
    q = O.DoublePriorityQueue()
    q.push(2.0, 2); q.push(3.0, 4); q.push(3.0, 3); q.push(1.0, 1)
    assert q.pop_last() == (3.0, 3) and len(q) == 3 and q.peek_first() == (1.0, 1) and q.peek_last_dist() == 3.0
    assert q.pop_last() == (3.0, 4) and len(q) == 2 and q.peek_last_dist() == 2.0
    assert q.pop_last() == (2.0, 2) and len(q) == 1 and q.peek_last_dist() == 1.0
    assert q.pop_last() == (1.0, 1) and len(q) == 0 and q.peek_first() is None


def test_bruteforce_language_test():
    # language-tests/tests/language/indexes/knn/bruteforce_knn_new_executor.surql:
    # pts:1..5 = [10,0],[2,0],[3,0],[100,0],[50,0]; point <|2,EUCLIDEAN|> [1,0] -> pts:2 (1f), pts:3 (2f)
    corpus = np.array([[10, 0], [2, 0], [3, 0], [100, 0], [50, 0]], np.float64)
    rows, dist = O.knn_topk(corpus, [1.0, 0.0], "euclidean", 2)
    assert list(rows) == [1, 2] and list(dist) == [1.0, 2.0]


def test_hnsw_language_test_bruteforce_leg():
    # hnsw_knn_new_executor.surql: pts = [1,2,3,4],[4,5,6,7],[8,9,10,11]; <|2,EUCLIDEAN|> [2,3,4,5]
    # -> [{dist: 2f, id: pts:1}, {dist: 4f, id: pts:2}]
    corpus = np.array([[1, 2, 3, 4], [4, 5, 6, 7], [8, 9, 10, 11]], np.float64)
    rows, dist = O.knn_topk(corpus, [2.0, 3.0, 4.0, 5.0], "euclidean", 2)
    assert list(rows) == [0, 1] and list(dist) == [2.0, 4.0]


def test_ties_keep_scan_order_and_strict_replacement():
    # knn_topk.rs:216-226: replace the worst only when STRICTLY closer; ties -> earlier row wins
    corpus = np.array([[3.0], [1.0], [3.0], [1.0], [2.0], [1.0]], np.float64)
    rows, dist = O.knn_topk(corpus, [0.0], "euclidean", 3)
    assert list(rows) == [1, 3, 5]
    rows, dist = O.knn_topk(corpus, [0.0], "euclidean", 4)
    assert list(rows) == [1, 3, 5, 4]
    rows, dist = O.knn_topk(corpus, [0.0], "euclidean", 5)
    assert list(rows) == [1, 3, 5, 4, 0]  # row 0 (3.0) kept over row 2 (3.0): later equal is not strictly closer


def test_skip_rows_and_k_larger_than_n():
    corpus = np.array([[1.0], [2.0], [3.0]], np.float64)
    rows, dist = O.knn_topk(corpus, [0.0], "euclidean", 10, skip=[0, 1, 0])
    assert list(rows) == [0, 2]
    rows, _ = O.knn_topk(corpus, [0.0], "euclidean", 0)
    assert rows.size == 0


def test_nan_distance_ordering_follows_number_cmp():
    # zero vector -> 0/0 -> NaN cosine distance.  Number::cmp uses total_cmp: a NEGATIVE NaN (what
    # x86-64 produces for 0.0/0.0) sorts before every number, so the reference returns that row first.
    corpus = np.array([[1.0, 0.0], [0.0, 0.0], [0.0, 1.0]], np.float64)
    rows, dist = O.knn_topk(corpus, [1.0, 0.0], "cosine", 3)
    assert np.isnan(dist[list(rows).index(1)])
    order = list(rows)
    assert order.index(0) < order.index(2)


def test_topk_matches_full_sort_random():
    rng = np.random.default_rng(11)
    corpus = rng.uniform(-20, 20, (2000, 16))  # reference generator range  idx/trees/knn.rs:635
    q = rng.uniform(-20, 20, 16)
    for metric, fn in (("cosine", O.f64_cosine_distance), ("euclidean", O.f64_euclidean)):
        d = np.array([fn(corpus[i], q) for i in range(corpus.shape[0])])
        order = np.lexsort((np.arange(d.size), d))[:10]
        rows, dist = O.knn_topk(corpus, q, metric, 10)
        assert list(rows) == list(order)
        assert list(dist) == list(d[order])


def test_batch_driver_equals_single():
    rng = np.random.default_rng(5)
    corpus = rng.uniform(-1, 1, (500, 32)).astype(np.float32)
    qs = rng.uniform(-1, 1, (7, 32))
    rows, dist = O.knn_topk_batch(corpus, qs, "cosine", 5, n_threads=3)
    for i in range(7):
        r, d = O.knn_topk(corpus, qs[i], "cosine", 5)
        assert list(rows[i]) == list(r) and list(dist[i]) == list(d)


def test_legacy_priority_list_keeps_the_k_nearest_with_an_arbitrary_boundary_tie():
    # idx/planner/knn.rs:11-106 -- with distinct distances it is the plain top-k; ties at the boundary are a free choice
    rng = np.random.default_rng(5)
    for trial in range(200):
        n, k = int(rng.integers(1, 40)), int(rng.integers(1, 8))
        d = [float(x) for x in rng.integers(0, 6 if trial % 2 else 1000, n)]
        must, tie, left = O.knn_priority_list(d, k)
        order = sorted(range(n), key=lambda r: (d[r], r))
        kth = d[order[min(k, n) - 1]]
        assert sorted(must) == sorted(r for r in range(n) if d[r] < kth or (d[r] == kth and not tie))[:len(must)]
        assert all(d[r] == kth for r in tie)
        assert len(must) + left == min(k, n)
        # KnnTopK's answer (ties by scan order) is one of the legacy path's possible answers
        topk = set(order[:k])
        assert set(must) <= topk and topk - set(must) <= set(tie)


def test_language_test_filtered_bruteforce_goldens():
    # language-tests/tests/language/indexes/knn/bruteforce_knn_with_filter_new_executor.surql: the predicate is
    # applied by the scan below KnnTopK, so the operator only ever sees the active rows
    pts = {1: ([10, 0], True), 2: ([2, 0], False), 3: ([3, 0], True), 4: ([100, 0], True), 5: ([50, 0], False)}
    active = [(i, p) for i, (p, a) in pts.items() if a]
    corpus = np.array([p for _, p in active], np.float64)
    rows, dist = O.knn_topk(corpus, np.array([1.0, 0.0]), "euclidean", 2)
    assert [active[int(r)][0] for r in rows] == [3, 1] and list(dist) == [2.0, 9.0]
    # ... multisource variant (bruteforce_knn_multisource_filter_new_executor.surql): Union of two tables, then Filter
    recs = [("pts:1", [10, 0], True), ("pts:2", [2, 0], False), ("pts:3", [3, 0], True),
            ("pts2:1", [1.5, 0], False), ("pts2:2", [4, 0], True)]
    act = [r for r in recs if r[2]]
    rows, dist = O.knn_topk(np.array([r[1] for r in act], np.float64), np.array([1.0, 0.0]), "euclidean", 2)
    assert [act[int(r)][0] for r in rows] == ["pts:3", "pts2:2"] and list(dist) == [2.0, 3.0]
