"""Staging decoders on the GPU (He / Hn values -> device arrays, through the C ABI) vs oracle/kvformats.py, and the
fused loader: an index loaded from raw KV values must answer exactly like the oracle walking the same graph."""
import numpy as np
import pytest

from oracle import kvformats as K
from oracle import pyoracle as O
from test_oracle_kvformats import KATS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from surrealdb_b200 import Context
    return Context(0)


def dev_buffer(n_rows, dim, dtype):
    import torch
    return torch.full((n_rows, dim), -7.0, dtype=torch.float32 if dtype == "F32" else torch.float64, device="cuda")


def test_reference_kat_values_decode(ctx):
    from surrealdb_b200 import staging as S
    import torch
    items = [(i, K.storekey_unescape(esc)[0]) for i, (_, _, esc) in enumerate(KATS)]
    for dt in ("F32", "F64"):
        out = dev_buffer(len(items), 3, dt)
        assert S.decode_vectors(ctx, items, 3, out.data_ptr(), len(items), dt) == 0
        torch.cuda.synchronize()
        assert out.cpu().numpy().tolist() == [[1.0, 2.0, 3.0]] * len(items)


@pytest.mark.parametrize("variant", K.VARIANTS)
@pytest.mark.parametrize("dim", [1, 2, 3, 250, 251, 768, 1536])
def test_vector_values_all_variants(ctx, variant, dim):
    from surrealdb_b200 import staging as S
    import torch
    rng = np.random.default_rng(dim * 5 + len(variant))
    n = 257
    if variant[0] == "F":
        src = rng.uniform(-20, 20, (n, dim)).astype(np.float32)  # f32-representable, so F64 -> F32 is exact too
        src[3, 0] = np.nan
        src[4, -1] = -0.0
    else:
        lim = 30000 if variant == "I16" else 1 << 22
        src = rng.integers(-lim, lim, (n, dim))
    ids = rng.permutation(n + 5)[:n]  # scattered destination rows, 5 rows never written
    items = [(int(ids[i]), K.ser_vector(variant, src[i])) for i in range(n)]
    for dt in ("F32", "F64"):
        out = dev_buffer(n + 5, dim, dt)
        present = torch.zeros(n + 5, dtype=torch.uint8, device="cuda")
        assert S.decode_vectors(ctx, items, dim, out.data_ptr(), n + 5, dt, present.data_ptr()) == 0
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        want = np.full((n + 5, dim), -7.0, got.dtype)
        want[ids] = src.astype(got.dtype)
        assert got.tobytes() == want.tobytes()
        pres = np.zeros(n + 5, np.uint8)
        pres[ids] = 1
        assert present.cpu().numpy().tolist() == pres.tolist()


def test_vector_values_rejects_and_inexact(ctx):
    from surrealdb_b200 import staging as S
    import torch
    good = K.ser_vector("F32", [1, 2, 3, 4])
    items = [(0, good), (1, K.ser_vector("F32", [1, 2, 3])),          # wrong dimension
             (2, b"\x02" + good[1:]),                                  # unknown revision
             (3, good[:1] + b"\x07" + good[2:]),                       # unknown variant
             (4, good[:-1]),                                           # truncated
             (9, good),                                                # destination row out of range
             (5, K.ser_vector("F64", [0.1, 1, 2, 3])),                 # not representable in f32
             (6, K.ser_vector("I64", [(1 << 40) + 1, 1, 2, 3])),       # idem
             (7, K.ser_vector("F64", [np.nan, np.inf, -np.inf, 0.5]))]  # fine
    out = dev_buffer(8, 4, "F32")
    assert S.decode_vectors(ctx, items, 4, out.data_ptr(), 8, "F32") == 7
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert got[0].tolist() == [1, 2, 3, 4]
    assert all((got[r] == -7).all() for r in (1, 2, 3, 4))
    assert np.isnan(got[7, 0]) and got[7, 1:].tolist() == [np.inf, -np.inf, 0.5]
    out64 = dev_buffer(8, 4, "F64")
    assert S.decode_vectors(ctx, items[6:], 4, out64.data_ptr(), 8, "F64") == 0  # 2^40+1 is exact in f64


def test_node_values(ctx):
    from surrealdb_b200 import staging as S
    rng = np.random.default_rng(3)
    n_elems = 5000
    lists = {}
    for node in rng.permutation(n_elems)[:3000]:
        deg = int(rng.integers(0, 33))
        lists[int(node)] = [int(x) for x in rng.integers(0, n_elems, deg)]          # duplicates happen
    lists.pop(19, None)
    lists[17] = [int(x) for x in rng.integers(0, 200, 300)]                          # > 32 neighbours, many repeats
    lists[18] = [1, 2, 3, n_elems, 4, 1 << 40, 5]                                    # edges to unknown elements
    items = [(node, K.node_to_val(lists[node])) for node in sorted(lists)]
    items.append((n_elems + 3, K.node_to_val([1])))                                  # node id out of range
    items.append((19, K.node_to_val([1, 2])[:-1]))                                   # truncated value
    row_ptr, col_idx, bad = S.decode_nodes(ctx, items, n_elems)
    assert bad == 4
    for node in range(n_elems):
        want = [e for e in K.load_node(K.node_to_val(lists[node])) if e < n_elems] if node in lists else []
        assert col_idx[row_ptr[node]:row_ptr[node + 1]].tolist() == want, node
    r0, c0, b0 = S.decode_nodes(ctx, [], 10)
    assert r0.tolist() == [0] * 11 and c0.size == 0 and b0 == 0


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_index_loaded_from_raw_kv_values_answers_like_the_oracle(ctx, metric):
    from surrealdb_b200.hnsw import HnswIndex
    rng = np.random.default_rng(11)
    dim = 48
    data = rng.uniform(-20, 20, (1200, dim)).astype(np.float32)
    h = O.Hnsw(dim, metric, m=8, efc=60, seed=5)
    for v in data:
        h.insert(v)
    g = h.export()
    n = data.shape[0]
    # what a KV range scan of the index would return (He, Hn per layer, Hs)
    he = [(e, K.ser_vector("F32", g["vectors"][e])) for e in range(n)]
    hn = []
    for rp, ci in g["layers"]:
        hn.append([(e, K.node_to_val(ci[rp[e]:rp[e + 1]])) for e in range(n) if rp[e + 1] > rp[e]])
    state = K.hnsw_state(int(g["entry_point"]), n, (n, 0), tuple((1, 0) for _ in g["layers"][1:]))
    idx = HnswIndex.from_kv(ctx, dim, state, he, hn, metric)
    assert idx.n_bad == 0 and idx.n == n
    queries = rng.uniform(-20, 20, (40, dim)).astype(np.float32)
    for k, ef in ((10, 40), (5, 5)):
        ids, dist, cnt, ctr = idx.search_graph(queries, k, ef, counters=True)
        for q in range(queries.shape[0]):
            oi, od, oc = O.hnsw_search_csr(g, queries[q], k, ef)
            assert list(ids[q, : cnt[q]]) == list(oi)
            assert dist[q, : cnt[q]].tobytes() == od.tobytes()
            assert (int(ctr[q, 0]), int(ctr[q, 1])) == oc
