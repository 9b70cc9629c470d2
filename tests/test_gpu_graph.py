"""CSR frontier expansion on the GPU (through the C ABI) vs the oracle and the reference's language tests."""
import json
import os
import re

import numpy as np
import pytest

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "graph_relations.json")))


@pytest.fixture(scope="module")
def ctx():
    from surrealdb_b200 import Context
    return Context(0)


@pytest.fixture(scope="module")
def store(ctx):
    from surrealdb_b200.graph import GraphStore
    return GraphStore(ctx, [(r["src"], r["edge_tb"], r["edge_id"], r["dst"]) for r in G["relations"]])


def fmt(names):
    return "[" + ", ".join(names) + "]"


def test_language_test_chains(store):
    checked = 0
    for f in ["traversal_multi_hop.surql", "traversal_forward.surql", "traversal_backward.surql"]:
        case = G["cases"][f]
        for stmt, res in zip(case["statements"], case["results"]):
            m = re.match(r"^(\w+:\w+)((?:(?:->|<-)\w+(?:->|<-)\w+)+);$", stmt)
            if not m:
                break
            hops = [("out" if a == "->" else "in", tb) for a, tb, _b, _t in re.findall(r"(->|<-)(\w+)(->|<-)(\w+)", m.group(2))]
            assert fmt(store.lookup([m.group(1)], hops)) == res, stmt
            checked += 1
    assert checked >= 9


def test_language_test_collect_and_recursion(store):
    c = G["cases"]["cycles_collect.surql"]["results"]
    assert fmt(store.collect("person:alice", "out", "knows", 1, 6, False)) == c[0]
    assert fmt(store.collect("person:alice", "out", "knows", 1, 6, True)) == c[1]
    c = G["cases"]["collect_min_depth.surql"]["results"]
    assert fmt(store.collect("person:alice", "out", "reports_to", 1, 256, False)) == c[0]
    assert fmt(store.collect("person:alice", "out", "reports_to", 3, 256, False)) == c[1]
    assert fmt(store.collect("person:alice", "out", "reports_to", 2, 3, False)) == c[2]
    assert fmt(store.collect("person:alice", "out", "reports_to", 2, 256, True)) == c[3]
    assert fmt(store.collect("person:alice", "out", "reports_to", 2, 2, False)) == c[4]
    c = G["cases"]["depth_fixed.surql"]["results"]
    for n in (1, 2, 3, 4):
        assert fmt(store.recurse("person:alice", "out", "reports_to", n, n)) == c[n - 1]
    assert fmt(store.recurse("person:alice", "out", "knows", 2, 2)) == c[4]
    c = G["cases"]["depth_range.surql"]["results"]
    assert fmt(store.recurse("person:alice", "out", "knows", 1, 3)) == c[4]


def rmat(n_nodes, n_edges, seed):
    rng = np.random.default_rng(seed)
    bits = int(np.ceil(np.log2(n_nodes)))
    src = np.zeros(n_edges, np.int64)
    dst = np.zeros(n_edges, np.int64)
    for b in range(bits):  # a,b,c,d = .57,.19,.19,.05
        r = rng.random(n_edges)
        src = (src << 1) | ((r >= 0.76) | ((r >= 0.57) & (r < 0.76) & False)).astype(np.int64)
        dst = (dst << 1) | (((r >= 0.57) & (r < 0.76)) | (r >= 0.95)).astype(np.int64)
    src %= n_nodes
    dst %= n_nodes
    order = np.lexsort((dst, src))  # integer edge ids assigned in (src,dst) order => KV order == this order
    src, dst = src[order], dst[order]
    rp = np.zeros(n_nodes + 1, np.uint64)
    np.add.at(rp, src + 1, 1)
    return np.cumsum(rp).astype(np.uint64), dst.astype(np.uint32)


@pytest.mark.parametrize("limit", [0, 3])
def test_rmat_multi_hop_equals_oracle(ctx, limit):
    from surrealdb_b200.graph import CsrGraph, expand
    rp, ci = rmat(1 << 16, 600_000, 3)
    g = CsrGraph(ctx, rp, ci)
    rng = np.random.default_rng(9)
    frontier = rng.integers(0, 1 << 16, 300).astype(np.uint32)
    frontier[10] = frontier[11]  # duplicates in the frontier are expanded again
    want = frontier
    for hop in range(3):
        want = O.graph_hop(rp, ci, want, limit)
        got = expand([g] * (hop + 1), frontier, limit)
        assert got.size == want.size and np.array_equal(got, want), (hop, got.size, want.size)
    assert want.size > 1000


def test_rmat_collect_equals_oracle(ctx):
    from surrealdb_b200.graph import CsrGraph, collect
    rp, ci = rmat(1 << 14, 100_000, 5)
    g = CsrGraph(ctx, rp, ci)
    for start, mn, mx, inc in ((5, 1, 0, False), (77, 2, 4, False), (123, 1, 3, True), (9000, 1, 0, True)):
        want = O.graph_collect(rp, ci, [start], mn, mx, inc)
        got = collect(g, [start], mn, mx, inc)
        assert np.array_equal(got, want), (start, mn, mx, inc, got.size, want.size)


def test_edge_cases(ctx):
    from surrealdb_b200 import SdbError
    from surrealdb_b200.graph import CsrGraph, expand
    rp = np.array([0, 3, 3, 5], np.uint64)
    ci = np.array([1, 2, 2, 0, 1], np.uint32)
    g = CsrGraph(ctx, rp, ci)
    assert list(expand([g], [0, 2, 0])) == [1, 2, 2, 0, 1, 1, 2, 2]
    assert list(expand([g], [0, 2, 0], 2)) == [1, 2, 0, 1, 1, 2]
    assert list(expand([g], [1])) == []
    assert list(expand([g], [])) == []
    assert list(expand([g, g], [1, 1, 1])) == []
    with pytest.raises(SdbError):
        expand([g], [3])
    # a slice of the output that spans thousands of zero-degree sources (global-search fallback in the kernel)
    n = 20000
    deg = np.zeros(n, np.int64)
    deg[::5000] = 3000
    rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint64)
    ci = (np.arange(rp[-1]) % n).astype(np.uint32)
    g = CsrGraph(ctx, rp, ci)
    fr = np.arange(n, dtype=np.uint32)
    assert np.array_equal(expand([g], fr), O.graph_hop(rp, ci, fr, 0))


def test_language_test_bidirectional(store):
    # language-tests/tests/language/graph/traversal_bidirectional.surql (results 0-2): `<->knows<->person`
    from test_oracle_graph import BIDIRECTIONAL
    for start, want in BIDIRECTIONAL.items():
        assert fmt(store.lookup([start], [("both", "knows")])) == want, start
    # two bidirectional hops in one fused call equal two single calls (multiset expansion keeps order and duplicates)
    one = store.lookup(["person:alice"], [("both", "knows")])
    two = store.lookup(["person:alice"], [("both", "knows"), ("both", "knows")])
    assert two == [x for n in one for x in store.lookup([n], [("both", "knows")])]
