"""Host-side mirror of the graph-scan operators on top of the C ABI.

  CsrGraph            sdb_graph: device CSR of one (direction, edge table)
  GraphStore          what the Rust shim builds from the `~` edge-pointer keys (key/graph/mod.rs:122-137):
                      per (direction, edge table) a CSR whose rows list, per source record, the targets in KV
                      key order of the connecting edge records
  GraphStore.lookup   LookupPart over a fused GraphEdgeScan chain (exec/parts/lookup.rs:139-170,
                      exec/planner/idiom.rs:161-193): multiset, order-preserving
  GraphStore.collect  `.{min..max+collect[+inclusive]}` (exec/operators/recursion/collect.rs:74-143)
  GraphStore.recurse  default `.{min..max}` recursion (exec/operators/recursion/default.rs:75-133)
  GraphEdgeScan       the operator itself: new(input, direction, edge_tables, output_mode, version).with_limit(n)
                      (exec/operators/scan/graph.rs:89-147) -- name(), attrs(), execute()
"""
import ctypes as C

import numpy as np

from . import _lib as L


class CsrGraph:
    def __init__(self, ctx, row_ptr, col_idx):
        self.ctx = ctx
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        ci = np.ascontiguousarray(col_idx, np.uint32)
        self.n_rows = rp.size - 1
        self.h = C.c_void_p()
        L.check(L.lib().sdb_graph_load_csr(ctx.h, self.n_rows, C.c_void_p(rp.ctypes.data),
                                           C.c_void_p(ci.ctypes.data) if ci.size else None, C.byref(self.h)))

    def close(self):
        if self.h:
            L.lib().sdb_graph_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CsrGraphShard(CsrGraph):
    """rows [row_lo, row_hi) of an n_rows_total-row CSR on this context's GPU (sdb_graph_load_csr_shard).  `row_ptr` /
    `col_idx` are the WHOLE graph's arrays (host): the slice is cut and rebased here.  expand / expand_device / collect
    on shard handles are collective over the context's communicator -- one thread or process per rank."""

    def __init__(self, ctx, row_ptr, col_idx, row_lo, row_hi):
        self.ctx = ctx
        rp_all = np.asarray(row_ptr, np.uint64)
        self.n_rows = rp_all.size - 1
        e0, e1 = int(rp_all[row_lo]), int(rp_all[row_hi])
        rp = np.ascontiguousarray(rp_all[row_lo:row_hi + 1] - np.uint64(e0))
        ci = np.ascontiguousarray(np.asarray(col_idx, np.uint32)[e0:e1])
        self.row_lo, self.row_hi = int(row_lo), int(row_hi)
        self.h = C.c_void_p()
        L.check(L.lib().sdb_graph_load_csr_shard(ctx.h, self.n_rows, self.row_lo, self.row_hi, C.c_void_p(rp.ctypes.data),
                                                 C.c_void_p(ci.ctypes.data) if ci.size else None, C.byref(self.h)))


def _take(out, n):
    if not out or n.value == 0:
        return np.zeros(0, np.uint32)
    arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint32)), shape=(n.value,)).copy()
    L.lib().sdb_free(out)
    return arr


def expand(hops, frontier, per_source_limit=0):
    """sdb_graph_expand: apply the CSR hops in order to the frontier (multiset semantics)."""
    fr = np.ascontiguousarray(frontier, np.uint32)
    arr = (C.c_void_p * len(hops))(*[g.h for g in hops])
    out, n = C.c_void_p(), C.c_uint64()
    L.check(L.lib().sdb_graph_expand(arr, len(hops), C.c_void_p(fr.ctypes.data) if fr.size else None, fr.size,
                                     int(per_source_limit), C.byref(out), C.byref(n)))
    return _take(out, n)


def expand_device(ctx, hops, d_frontier, n_frontier, per_source_limit=0):
    """sdb_graph_expand_device: frontier and result stay in HBM. -> (device pointer (int), count); free with
    device_free(ctx, ptr)."""
    arr = (C.c_void_p * len(hops))(*[g.h for g in hops])
    out, n = C.c_void_p(), C.c_uint64()
    L.check(L.lib().sdb_graph_expand_device(arr, len(hops), C.c_void_p(d_frontier), int(n_frontier), int(per_source_limit),
                                            C.byref(out), C.byref(n)))
    return (out.value or 0), n.value


def device_free(ctx, ptr):
    if ptr:
        L.lib().sdb_device_free(ctx.h, C.c_void_p(ptr))


def collect(graph, start, min_depth=1, max_depth=0, inclusive=False):
    st = np.ascontiguousarray(start, np.uint32)
    out, n = C.c_void_p(), C.c_uint64()
    L.check(L.lib().sdb_graph_collect(graph.h, C.c_void_p(st.ctypes.data) if st.size else None, st.size, int(min_depth),
                                      int(max_depth), int(bool(inclusive)), C.byref(out), C.byref(n)))
    return _take(out, n)


def _key_order(rid_key):
    """storekey order of a RecordIdKey (val/record_id.rs:181-192): numbers (numeric) before strings (bytes)"""
    if isinstance(rid_key, (int, np.integer)):
        return (0, int(rid_key), b"")
    return (1, 0, str(rid_key).encode())


class GraphStore:
    """CSR snapshots of a set of RELATE edges: relations = iterable of (src, edge_table, edge_id, dst) with
    record ids like 'person:alice'."""

    def __init__(self, ctx, relations):
        self.ctx = ctx
        rel = list(relations)
        self.names = sorted({r[0] for r in rel} | {r[3] for r in rel})
        self.idx = {n: i for i, n in enumerate(self.names)}
        self._rel = rel
        self._csr = {}

    def csr_arrays(self, edge_table, direction):
        """(row_ptr, col_idx) of one `node <dir> edge_table <dir> node` step, neighbours in the order the reference's
        KV scan yields them: per source the graph keys sort by (direction, edge table, edge record key)
        (key/graph/mod.rs:122-137).  edge_table=None is the `?` wildcard (all edge tables, scan/graph.rs:303-311); a tuple of
        names = one key range per table, scanned in the listed order (scan/graph.rs:312-324).  direction: 'out' (->), 'in' (<-) or 'both' (<->): GraphEdgeScan scans In, then
        Out (exec/operators/scan/graph.rs:203-207), and the second `<->` of the pair yields both endpoints of every
        edge record, In pointer (the edge's source node) first."""
        adj = [[] for _ in self.names]
        for src, tb, eid, dst in self._rel:
            if edge_table is None:  # the `?` wildcard: one range over every edge table, i.e. key order by table name
                tkey = tb.encode()
            elif isinstance(edge_table, (tuple, list)):  # one range per listed table, scanned in the listed order
                if tb not in edge_table:
                    continue
                tkey = list(edge_table).index(tb)
            else:
                if tb != edge_table:
                    continue
                tkey = 0
            ek = (tkey, _key_order(eid))  # `ft` (the edge table) sorts before `fk` (the edge record key)
            s, d = self.idx[src], self.idx[dst]
            if direction == "out":
                adj[s].append(((1, ek, 0), d))
            elif direction == "in":
                adj[d].append(((0, ek, 0), s))
            elif direction == "both":
                adj[d].append(((0, ek, 0), s))  # edges pointing at d: [source, d]
                adj[d].append(((0, ek, 1), d))
                adj[s].append(((1, ek, 0), s))  # edges leaving s: [s, target]
                adj[s].append(((1, ek, 1), d))
            else:
                raise ValueError(f"direction {direction!r}: expected 'out', 'in' or 'both'")
        rp, ci = [0], []
        for a in adj:
            a.sort()
            ci += [t for _, t in a]
            rp.append(len(ci))
        return np.asarray(rp, np.uint64), np.asarray(ci, np.uint32)

    def csr(self, edge_table, direction):
        key = (tuple(edge_table) if isinstance(edge_table, list) else edge_table, direction)
        if key not in self._csr:
            rp, ci = self.csr_arrays(edge_table, direction)
            self._csr[key] = CsrGraph(self.ctx, rp, ci)
        return self._csr[key]

    def expand_snapshot(self, edge_table, direction, frontier, per_source_limit=0):
        """one GraphEdgeScan step over the snapshot of (edge tables, direction) on the GPU"""
        return expand([self.csr(edge_table, direction)], frontier, per_source_limit)

    def ids(self, names):
        return np.asarray([self.idx[n] for n in names], np.uint32)

    def to_names(self, arr):
        return [self.names[int(i)] for i in arr]

    def lookup(self, start, hops, limit=0):
        """start: record ids; hops: [(direction 'out'|'in', edge_table), ...] -> record ids (order + duplicates
        as the reference returns them)"""
        return self.to_names(expand([self.csr(tb, d) for d, tb in hops], self.ids(start), limit))

    def collect(self, start, direction, edge_table, min_depth=1, max_depth=0, inclusive=False):
        return self.to_names(collect(self.csr(edge_table, direction), self.ids([start]), min_depth, max_depth, inclusive))

    def recurse(self, start, direction, edge_table, min_depth, max_depth):
        """default recursion: repeat the hop until the bound, a dead end or a fixed point"""
        g = self.csr(edge_table, direction)
        cur = self.ids([start])
        depth = 0
        while depth < max_depth:
            nxt = expand([g], cur)
            depth += 1
            if nxt.size == 0 or (nxt.size == cur.size and np.array_equal(nxt, cur)):
                return self.to_names(cur) if depth > min_depth else None
            cur = nxt
        return self.to_names(cur) if depth >= min_depth else None


class GraphEdgeScan:
    """Mirror of the reference operator (exec/operators/scan/graph.rs:89-147):
    `GraphEdgeScan::new(input, direction, edge_tables, output_mode, version)` + `.with_limit(n)`, `name()`, `attrs()`,
    `execute()`.  `input` yields the source record ids (what the child operator streams); the result is the target
    record id of every matching edge pointer, per source in KV key order, duplicates kept -- served by one
    sdb_graph_expand over the CSR snapshot of (direction, edge tables).  Range bounds on an edge table, `version`
    (time travel) and output modes other than TargetId are the cases INTEGRATION.md leaves on the KV path."""

    DIRECTIONS = {"->": "out", "<-": "in", "<->": "both"}

    def __init__(self, input, direction, edge_tables, output_mode="TargetId", version=None, store=None):
        if direction not in self.DIRECTIONS:
            raise L.SdbError(L.SDB_EUNSUPPORTED, f"direction {direction!r} is not served by the GPU snapshot")
        if output_mode != "TargetId" or version is not None:
            raise L.SdbError(L.SDB_EUNSUPPORTED, "only GraphScanOutput::TargetId without VERSION is served by the GPU snapshot")
        self.input, self.direction, self.edge_tables = input, direction, list(edge_tables)
        self.output_mode, self.version, self.limit, self.store = output_mode, version, None, store

    def with_limit(self, limit):
        self.limit = int(limit)
        return self

    def name(self):
        return "GraphEdgeScan"

    def attrs(self):
        a = [("direction", self.direction), ("tables", ", ".join(self.edge_tables) if self.edge_tables else "*"),
             ("output", self.output_mode)]
        if self.limit is not None:
            a.append(("limit", str(self.limit)))
        return a

    def _snapshot(self):
        tables = None if not self.edge_tables else (self.edge_tables[0] if len(self.edge_tables) == 1 else tuple(self.edge_tables))
        return tables, self.DIRECTIONS[self.direction]

    def execute(self):
        tables, d = self._snapshot()
        sources = list(self.input)
        unknown = [s for s in sources if s not in self.store.idx]  # a record without edges has no graph keys: no output
        frontier = self.store.ids([s for s in sources if s in self.store.idx])
        del unknown
        if self.limit == 0:
            # with_limit(0): the per-source key stream is opened with limit Some(0) and yields nothing
            # (scan/graph.rs:238 -> kvs/scanner.rs:160-172: ScanLimit::Count(min(batch, 0))); None = unlimited
            return []
        return self.store.to_names(self.store.expand_snapshot(tables, d, frontier, self.limit or 0))
