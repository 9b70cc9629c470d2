"""Round-2 brute-force machinery on the GPU, all through the C ABI and all against the CPU oracle:
  * the streaming tensor-core screen (one launch, in-kernel threshold refinement) gives the same bits as the
    multi-pass schedule and as the oracle -- on spread-out, clustered, sorted and duplicate-heavy corpora;
  * candidate sets that adapt to the data (thousands of rows inside a tight cluster) go through the chunked final
    sort and, when they overflow the lists, up the precision ladder -- never silently wrong;
  * the adversarial bf16 case of ADVICE r1 (every component at a rounding midpoint);
  * asynchronous batches (submit / wait, device and host buffers, out-of-order completion);
  * the sharded entry points on one rank (block layout, header, merge) and, with >= 2 GPUs, one process driving
    two devices (sdb_ctx_create_multi + sdb_knn_sharded_multi) and two contexts on two devices.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from surrealdb_b200 import Context
    return Context(0)


def make_col(ctx, corpus, metric, screen=None, streaming=True):
    from surrealdb_b200 import VectorColumn
    col = VectorColumn(ctx, corpus.shape[1], metric, "F32", capacity=max(1, corpus.shape[0]))
    col.append(corpus)
    col.finalize()
    if screen:
        col.set_screen(screen)
    col.set_schedule(streaming)
    return col


def check_queries(col, corpus, queries, metric, k, which):
    rows, dist, cnt = col.knn(queries, k)
    for q in which:
        r, d = O.knn_topk(corpus, queries[q], metric.lower(), k)
        assert cnt[q] == r.size, (q, cnt[q], r.size)
        assert list(rows[q, : cnt[q]]) == list(r), (q, rows[q], r)
        assert dist[q, : cnt[q]].tobytes() == d.tobytes(), (q, dist[q], d)
    return rows, dist, cnt


def clustered(rng, n, dim, n_clusters, sigma):
    cent = rng.normal(0, 1, (n_clusters, dim))
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    lab = rng.integers(0, n_clusters, n)
    x = cent[lab] + rng.normal(0, sigma / np.sqrt(dim), (n, dim))
    return x.astype(np.float32), cent, lab


@pytest.mark.parametrize("metric,screen", [("COSINE", "TC_INT8"), ("COSINE", "TC_BF16"), ("EUCLIDEAN", "TC_BF16")])
def test_streaming_equals_multipass_equals_oracle(ctx, metric, screen):
    rng = np.random.default_rng(len(metric) + len(screen))
    n, dim, nq = 60_000, 96, 300  # 235 tiles: a scored sample of 16 + one streaming launch over the other 219
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    queries = rng.uniform(-1, 1, (nq, dim))
    col_s = make_col(ctx, corpus, metric, screen, streaming=True)
    col_p = make_col(ctx, corpus, metric, screen, streaming=False)
    for k in (1, 10, 100):
        rs, ds, cs = check_queries(col_s, corpus, queries, metric, k, (0, 7, 150, 299))
        st = col_s.stats()
        assert st["n_passes"] <= 2 and st["n_fallback"] == 0, st
        rp, dp, cp = col_p.knn(queries, k)
        assert col_p.stats()["n_passes"] >= 3
        assert rs.tobytes() == rp.tobytes() and ds.tobytes() == dp.tobytes() and cs.tobytes() == cp.tobytes()


def test_streaming_on_sorted_and_clustered_corpora(ctx):
    # rows ordered by cluster (the streaming launch meets whole clusters at once) and queries inside clusters:
    # the candidate set of a query is its whole cluster neighbourhood, far more than k
    rng = np.random.default_rng(99)
    n, dim, nq, k = 80_000, 128, 64, 10
    x, cent, lab = clustered(rng, n, dim, 40, 0.15)
    order = np.argsort(lab, kind="stable")
    corpus = np.ascontiguousarray(x[order])
    qc = rng.integers(0, 40, nq)
    queries = (cent[qc] + rng.normal(0, 0.15 / np.sqrt(dim), (nq, dim))).astype(np.float64)
    for screen in ("TC_INT8", "TC_BF16"):
        col = make_col(ctx, corpus, "COSINE", screen)
        check_queries(col, corpus, queries, "COSINE", k, (0, 13, 63))
        st = col.stats()
        assert st["n_fallback"] <= 2 + nq // 64, st  # big candidate sets are re-ranked, not sent to the exact kernel
        assert st["n_survivors"] > 200 * nq, st  # ... and they ARE big here (2000 rows per cluster); the f32 stage
        assert st["n_candidates"] < 200, st      # then narrows them to a few near-ties before the f64 re-rank
    # the same through the multi-pass schedule
    col = make_col(ctx, corpus, "COSINE", "TC_BF16", streaming=False)
    check_queries(col, corpus, queries, "COSINE", k, (5, 40))


def test_candidate_overflow_climbs_the_ladder(ctx):
    # one dense cluster of 9000 rows: with the int8 screen's margin every one of them is a candidate (> 4096 slots)
    rng = np.random.default_rng(5)
    dim, k = 64, 10
    x, cent, lab = clustered(rng, 9000, dim, 1, 0.05)
    far = rng.uniform(-1, 1, (6000, dim)).astype(np.float32)
    corpus = np.concatenate([x, far])[rng.permutation(15000)]
    queries = (cent[[0] * 20] + rng.normal(0, 0.05 / np.sqrt(dim), (20, dim))).astype(np.float64)
    col = make_col(ctx, corpus, "COSINE")  # AUTO
    check_queries(col, corpus, queries, "COSINE", k, (0, 9, 19))
    # whatever rung or fallback it took, the second batch starts where the first one settled
    rows2, dist2, cnt2 = col.knn(queries, k)
    rows1, dist1, cnt1 = col.knn(queries, k)
    assert rows1.tobytes() == rows2.tobytes() and dist1.tobytes() == dist2.tobytes()


def test_bf16_rounding_midpoints_do_not_break_the_proof(ctx):
    # every component sits exactly between two bf16 values, so BOTH operands round the same way and the
    # screened similarity of the true neighbours is biased by ~2^-7; the error bound must account for both roundings.
    rng = np.random.default_rng(3)
    dim, n = 64, 30_000
    mid = np.float32(1.0 + 2.0 ** -8)  # halfway between bf16(1.0) and the next bf16 value
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    corpus[:2000] = mid * np.sign(corpus[:2000])
    corpus[:2000] *= (1 + rng.integers(0, 4, (2000, dim)) * np.float32(2.0 ** -12))  # many near-ties, all near midpoints
    queries = (mid * np.sign(corpus[:8])).astype(np.float64)
    for screen in ("TC_BF16", "TC_INT8"):
        col = make_col(ctx, corpus, "COSINE", screen)
        check_queries(col, corpus, queries, "COSINE", 10, range(8))
    col = make_col(ctx, corpus, "EUCLIDEAN", "TC_BF16")
    check_queries(col, corpus, queries, "EUCLIDEAN", 10, range(8))


def test_outlier_rows_do_not_dictate_the_int8_scale(ctx):
    # a few rows with one dominant component: they become special rows (ranked exactly on every query) and the int8
    # copy is scaled for the others, so AUTO keeps the int8 screen; an outlier that IS the nearest neighbour is found
    rng = np.random.default_rng(23)
    n, dim, k = 20_000, 256, 10
    corpus = rng.normal(0, 1, (n, dim)).astype(np.float32)
    out_rows = [7, 1234, 19_999]
    for r in out_rows:
        corpus[r, 5] *= 80.0
    queries = rng.normal(0, 1, (6, dim))
    queries[0] = corpus[1234].astype(np.float64) * 1.01  # nearest neighbour = an outlier row
    col = make_col(ctx, corpus, "COSINE")  # AUTO
    check_queries(col, corpus, queries, "COSINE", k, range(6))
    st = col.stats()
    assert st["screen_used"] == 4 and st["n_special_rows"] == len(out_rows) and st["n_fallback"] == 0, st


def test_tombstones_without_refinalize(ctx):
    # sdb_corpus_remove: rows disappear from every later search (all screens and the exact kernel) without a re-finalize
    rng = np.random.default_rng(29)
    n, dim, k = 30_000, 48, 10
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    corpus[11] = 0.0  # a special row (zero norm)
    queries = rng.uniform(-1, 1, (4, dim))
    for metric, screens in (("COSINE", ("TC_INT8", "TC_BF16", "SIMT_F32", "NONE_EXACT")), ("EUCLIDEAN", ("TC_BF16", "NONE_EXACT"))):
        col = make_col(ctx, corpus, metric)
        rows0, _, _ = col.knn(queries, k)
        gone = np.unique(np.concatenate([rows0[:, :3].ravel(), [11, 0, n - 1]])).astype(np.uint64)
        col.remove(gone)
        skip = np.zeros(n, np.uint8)
        skip[gone] = 1
        for screen in screens:
            col.set_screen(screen)
            rows, dist, cnt = col.knn(queries, k)
            for q in range(4):
                r, d = O.knn_topk(corpus, queries[q], metric.lower(), k, skip=skip)
                assert list(rows[q, : cnt[q]]) == list(r), (metric, screen, q)
                assert dist[q, : cnt[q]].tobytes() == d.tobytes()
        # a later skip mask does not resurrect the removed rows, and finalize keeps them out
        extra = np.zeros(n, np.uint8)
        extra[100:200] = 1
        col.set_skip(extra)
        col.finalize()
        col.set_screen("AUTO")
        both = skip | extra
        rows, dist, cnt = col.knn(queries, k)
        for q in range(4):
            r, d = O.knn_topk(corpus, queries[q], metric.lower(), k, skip=both)
            assert list(rows[q, : cnt[q]]) == list(r) and dist[q, : cnt[q]].tobytes() == d.tobytes()


def test_filtered_and_tiny_samples(ctx):
    # a skip mask that leaves fewer than k valid rows in the scored sample: the thresholds start at -inf and the
    # histograms are seeded from the score range instead
    from surrealdb_b200 import VectorColumn
    rng = np.random.default_rng(17)
    n, dim, k = 40_000, 32, 10
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    queries = rng.uniform(-1, 1, (5, dim))
    skip = (rng.uniform(0, 1, n) > 0.002).astype(np.uint8)  # ~80 valid rows
    for screen in ("TC_INT8", "TC_BF16"):
        col = VectorColumn(ctx, dim, "COSINE", "F32", capacity=n)
        col.append(corpus)
        col.set_skip(skip)
        col.finalize()
        col.set_screen(screen)
        rows, dist, cnt = col.knn(queries, k)
        for q in range(5):
            r, d = O.knn_topk(corpus, queries[q], "cosine", k, skip=skip)
            assert list(rows[q, : cnt[q]]) == list(r) and dist[q, : cnt[q]].tobytes() == d.tobytes()


def _dev_arrays(torch, dev, nq, k):
    return (torch.zeros((nq, k), dtype=torch.int64, device=dev), torch.zeros((nq, k), dtype=torch.float64, device=dev),
            torch.zeros((nq,), dtype=torch.int32, device=dev))


def test_async_batches_device_and_host(ctx):
    import torch
    rng = np.random.default_rng(31)
    n, dim, nq, k = 50_000, 64, 200, 10
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    col = make_col(ctx, corpus, "COSINE")
    dev = torch.device("cuda", 0)
    batches = [rng.uniform(-1, 1, (nq, dim)) for _ in range(4)]
    want = [col.knn(b, k) for b in batches]
    # device buffers, four batches in flight, completed out of order
    qd = [torch.from_numpy(b).to(dev) for b in batches]
    outs = [_dev_arrays(torch, dev, nq, k) for _ in batches]
    torch.cuda.synchronize()
    tickets = [col.submit_device(qd[i].data_ptr(), nq, k, 1000, outs[i][0].data_ptr(), outs[i][1].data_ptr(),
                                 outs[i][2].data_ptr()) for i in range(4)]
    with pytest.raises(Exception, match="in flight"):
        col.submit_device(qd[0].data_ptr(), nq, k, 0, outs[0][0].data_ptr(), outs[0][1].data_ptr(), outs[0][2].data_ptr())
    for i in (2, 0, 3, 1):
        col.wait(tickets[i])
    with pytest.raises(Exception, match="ticket"):
        col.wait(tickets[0])
    for i in range(4):
        assert (outs[i][0].cpu().numpy() - 1000).astype(np.uint64).tobytes() == want[i][0].tobytes()
        assert outs[i][1].cpu().numpy().tobytes() == want[i][1].tobytes()
        assert (outs[i][2].cpu().numpy() == k).all()
    # pinned host buffers: H2D on the copy stream, D2H behind the batch
    hq = [torch.from_numpy(b).pin_memory() for b in batches]
    hr = [torch.zeros((nq, k), dtype=torch.int64).pin_memory() for _ in batches]
    hd = [torch.zeros((nq, k), dtype=torch.float64).pin_memory() for _ in batches]
    hc = [torch.zeros((nq,), dtype=torch.int32).pin_memory() for _ in batches]
    tickets = [col.submit_host(hq[i].data_ptr(), nq, k, hr[i].data_ptr(), hd[i].data_ptr(), hc[i].data_ptr())
               for i in range(4)]
    for t in tickets:
        col.wait(t)
    for i in range(4):
        assert hr[i].numpy().astype(np.uint64).tobytes() == want[i][0].tobytes()
        assert hd[i].numpy().tobytes() == want[i][1].tobytes()
    # a batch that needs host-side repairs (zero query -> exact path) while another one is in flight behind it
    special = batches[0].copy()
    special[3] = 0.0
    sq = torch.from_numpy(special).to(dev)
    torch.cuda.synchronize()
    t0 = col.submit_device(sq.data_ptr(), nq, k, 0, outs[0][0].data_ptr(), outs[0][1].data_ptr(), outs[0][2].data_ptr())
    t1 = col.submit_device(qd[1].data_ptr(), nq, k, 0, outs[1][0].data_ptr(), outs[1][1].data_ptr(), outs[1][2].data_ptr())
    col.wait(t0)
    assert col.stats()["n_fallback"] == 1
    col.wait(t1)
    r, d = O.knn_topk(corpus, special[3], "cosine", k)
    assert list(outs[0][0].cpu().numpy()[3]) == list(r) and outs[0][1].cpu().numpy()[3].tobytes() == d.tobytes()
    assert outs[1][0].cpu().numpy().astype(np.uint64).tobytes() == want[1][0].tobytes()


def test_sharded_entry_points_on_one_rank(ctx):
    # nranks = 1: the block / header / merge path without NCCL; row_base shifts the returned ids
    import torch
    rng = np.random.default_rng(41)
    n, dim, nq, k = 30_000, 48, 50, 10
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    queries = rng.uniform(-1, 1, (nq, dim))
    queries[7] = 0.0  # repaired on the host side -> the repair round (second gather + merge) runs
    col = make_col(ctx, corpus, "COSINE")
    col.set_row_base(5_000_000_000)
    hq = torch.from_numpy(queries).pin_memory()
    hr = torch.zeros((nq, k), dtype=torch.int64).pin_memory()
    hd = torch.zeros((nq, k), dtype=torch.float64).pin_memory()
    hc = torch.zeros((nq,), dtype=torch.int32).pin_memory()
    t = col.sharded_submit_host(hq.data_ptr(), nq, k, hr.data_ptr(), hd.data_ptr(), hc.data_ptr())
    col.sharded_wait(t)
    for q in (0, 7, 49):
        r, d = O.knn_topk(corpus, queries[q], "cosine", k)
        assert list(hr.numpy()[q].astype(np.uint64) - np.uint64(5_000_000_000)) == list(r)
        assert hd.numpy()[q].tobytes() == d.tobytes()


def test_context_cancellation_flag():
    # sdb_ctx_cancel: brute force, HNSW walk and graph expansion all observe the context's flag (ctx.is_done() polls)
    from surrealdb_b200 import Context
    from surrealdb_b200.graph import CsrGraph, collect, expand
    from surrealdb_b200.hnsw import HnswIndex
    c2 = Context(0)
    rng = np.random.default_rng(71)
    corpus = rng.uniform(-1, 1, (5000, 16)).astype(np.float32)
    col = make_col(c2, corpus, "COSINE")
    q = rng.uniform(-1, 1, (3, 16))
    rp, ci = _rmat(rng, 10, 6000)
    g = CsrGraph(c2, rp, ci)
    layers = [(np.arange(5001, dtype=np.uint64) * 0, np.zeros(0, np.uint32))]
    layers[0] = (np.concatenate([[0], np.cumsum(np.full(5000, 2))]).astype(np.uint64),
                 np.stack([(np.arange(5000) + 1) % 5000, (np.arange(5000) + 7) % 5000], 1).ravel().astype(np.uint32))
    idx = HnswIndex(c2, corpus, layers, 0, "euclidean")
    want = col.knn(q, 5)
    c2.cancel()
    for call in (lambda: col.knn(q, 5), lambda: expand([g, g], np.arange(50, dtype=np.uint32), 0),
                 lambda: collect(g, np.arange(2, dtype=np.uint32), 1, 3, False),
                 lambda: idx.search_graph(corpus[:4], 3, 8)):
        with pytest.raises(Exception, match="SDB_ECANCELLED"):
            call()
    c2.cancel_reset()
    got = col.knn(q, 5)
    assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes()
    assert expand([g], np.arange(5, dtype=np.uint32), 0).size >= 0
    assert idx.search_graph(corpus[:4], 3, 8)[2].tolist() == [3, 3, 3, 3]


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def test_one_process_two_gpus():
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    from surrealdb_b200 import Context, VectorColumn
    from surrealdb_b200.engine import knn_sharded_multi
    from surrealdb_b200.sharding import shard_range
    rng = np.random.default_rng(51)
    n, dim, nq, k = 40_000, 64, 33, 10
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    corpus[30000:30004] = corpus[100:104]  # cross-shard exact ties resolve by global row
    queries = rng.uniform(-1, 1, (nq, dim))
    queries[5] = 0.0
    ctxs = Context.create_multi([0, 1])
    assert ctxs[0].comm_size() == 2
    shards = []
    for r, c in enumerate(ctxs):
        base, n_local = shard_range(n, 2, r)
        col = VectorColumn(c, dim, "COSINE", "F32", capacity=n_local)
        col.append(corpus[base:base + n_local])
        col.finalize()
        col.set_row_base(base)
        shards.append(col)
    for screen in ("TC_INT8", "TC_BF16"):  # both screens on BOTH devices of this process (per-device kernel attributes)
        for s in shards:
            s.set_screen(screen)
        rows, dist, cnt = knn_sharded_multi(shards, queries, k)
        for q in range(nq):
            r, d = O.knn_topk(corpus, queries[q], "cosine", k)
            assert list(rows[q]) == list(r) and dist[q].tobytes() == d.tobytes(), (screen, q)
    # two independent contexts on two devices in one process (no communicator): k = 100 forces the long lists
    c1 = Context(1)
    col1 = VectorColumn(c1, dim, "COSINE", "F32", capacity=n)
    col1.append(corpus)
    col1.finalize()
    rows, dist, cnt = col1.knn(queries[:3], 100)
    for q in range(3):
        r, d = O.knn_topk(corpus, queries[q], "cosine", 100)
        assert list(rows[q]) == list(r) and dist[q].tobytes() == d.tobytes()


def _rmat(rng, bits, n_edges):
    src = np.zeros(n_edges, np.int64)
    dst = np.zeros(n_edges, np.int64)
    for _ in range(bits):
        r = rng.uniform(0, 1, n_edges)
        src = (src << 1) | (r >= 0.76)
        dst = (dst << 1) | (((r >= 0.57) & (r < 0.76)) | (r >= 0.95))
    key = np.unique(src * (1 << bits) + dst)
    src, dst = key >> bits, key & ((1 << bits) - 1)
    rp = np.zeros((1 << bits) + 1, np.uint64)
    rp[1:] = np.cumsum(np.bincount(src, minlength=1 << bits))
    return rp, dst.astype(np.uint32)


def test_sharded_graph_on_one_rank(ctx):
    # a shard covering a sub-range of the rows on a single rank: sources outside the range expand to nothing here
    from surrealdb_b200.graph import CsrGraph, CsrGraphShard, expand
    rng = np.random.default_rng(61)
    rp, ci = _rmat(rng, 12, 40_000)
    whole = CsrGraph(ctx, rp, ci)
    full_shard = CsrGraphShard(ctx, rp, ci, 0, rp.size - 1)
    frontier = rng.integers(0, rp.size - 1, 500).astype(np.uint32)
    want = expand([whole, whole], frontier, 7)
    assert np.array_equal(expand([full_shard, full_shard], frontier, 7), want)
    assert np.array_equal(want, O.graph_hop(rp, ci, O.graph_hop(rp, ci, frontier, 7), 7))


def test_sharded_graph_two_gpus_threads():
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    import threading
    from surrealdb_b200 import Context
    from surrealdb_b200.graph import CsrGraphShard, collect, expand
    rng = np.random.default_rng(62)
    rp, ci = _rmat(rng, 13, 90_000)
    n = rp.size - 1
    frontier = rng.integers(0, n, 300).astype(np.uint32)
    want = frontier
    for _ in range(3):
        want = O.graph_hop(rp, ci, want, 5)
    want_collect = O.graph_collect(rp, ci, frontier[:2], 1, 4, False) if hasattr(O, "graph_collect") else None
    ctxs = Context.create_multi([0, 1])
    cut = n // 3  # uneven shards
    out, errs = [None, None], []

    def run(r):
        try:
            lo, hi = (0, cut) if r == 0 else (cut, n)
            g = CsrGraphShard(ctxs[r], rp, ci, lo, hi)
            res = expand([g, g, g], frontier, 5)
            col = collect(g, frontier[:2], 1, 4, False)
            out[r] = (res, col)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert not errs, errs
    for r in range(2):
        assert np.array_equal(out[r][0], want)
        if want_collect is not None:
            assert np.array_equal(out[r][1], want_collect)
    assert np.array_equal(out[0][1], out[1][1])


def test_failed_queries_climb_the_rungs_as_a_small_batch(ctx):
    # Two queries sit on a crowd of 6000 near-duplicates: all of them fall inside the int8 margin, the 4096-entry lists
    # overflow and the proof cannot succeed on the batch's rung.  Only THESE queries are re-screened on the finer rungs
    # (a small batch of their own) -- the rest of the batch is untouched, nothing reaches the exact kernel, and the
    # answers are the oracle's bit for bit.
    rng = np.random.default_rng(77)
    n, dim, nq, k = 80_000, 128, 256, 10
    corpus = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    center = rng.uniform(-1, 1, dim)
    corpus[1000:7000] = (center[None, :] + rng.normal(0, 2e-3, (6000, dim))).astype(np.float32)
    queries = rng.uniform(-1, 1, (nq, dim))
    queries[5] = center
    queries[77] = center + rng.normal(0, 1e-3, dim)
    col = make_col(ctx, corpus, "COSINE", screen="TC_INT8")
    rows, dist, cnt = check_queries(col, corpus, queries, "COSINE", k, (0, 5, 77, 100, 255))
    st = col.stats()
    assert st["screen_used"] == 4, st
    assert st["n_repaired"] + st["n_fallback"] >= 2, st  # the two crowd queries cannot be proven on the batch's rung
    # the same through the asynchronous entry points with another batch in flight
    qa, qb = np.ascontiguousarray(queries), np.ascontiguousarray(queries[::-1])
    outs = [(np.zeros((nq, k), np.uint64), np.zeros((nq, k), np.float64), np.zeros(nq, np.uint32)) for _ in range(2)]
    t1 = col.submit_host(qa.ctypes.data, nq, k, outs[0][0].ctypes.data, outs[0][1].ctypes.data, outs[0][2].ctypes.data)
    t2 = col.submit_host(qb.ctypes.data, nq, k, outs[1][0].ctypes.data, outs[1][1].ctypes.data, outs[1][2].ctypes.data)
    col.wait(t1)
    col.wait(t2)
    assert outs[0][0].tobytes() == rows.astype(np.uint64).tobytes() and outs[0][1].tobytes() == dist.tobytes()
    assert outs[1][0][::-1].tobytes() == rows.astype(np.uint64).tobytes()
