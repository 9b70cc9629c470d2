"""Host-side mirror of the HNSW search path on top of the C ABI.

  HnswIndex.load(...)                 what the Rust shim hands over after HnswFlavor::check_state
                                      (idx/trees/hnsw/mod.rs:187-224): element vectors + per-layer adjacency
  HnswIndex.search_graph(q, k, ef)    Hnsw::knn_search (hnsw/mod.rs:459-482) -> [(dist, element)] ascending
  HnswIndex.knn_search(q, k, ef)      HnswIndex::knn_search (hnsw/index.rs:270-335) without pendings/filter:
                                      element -> docs expansion through KnnResultBuilder semantics
                                      (idx/trees/knn.rs:363-437): final order (distance, doc id), <= k
"""
import ctypes as C

import numpy as np

from . import _lib as L


class HnswIndex:
    def __init__(self, ctx, vectors, layers, entry_point, metric="EUCLIDEAN", elem_docs=None):
        """vectors (n, dim) f32; layers = [(row_ptr u64[n+1], col_idx u32[e]), ...] layer 0 first;
        elem_docs: optional list of doc-id lists per element (identical vectors share one element:
        hnsw/docs.rs:161-176); default = one doc per element with the same id."""
        vec = np.ascontiguousarray(vectors, np.float32)
        self.n, self.dim = vec.shape
        self.ctx, self.metric = ctx, metric.upper()
        self.elem_docs = elem_docs
        nl = len(layers)
        rps = [np.ascontiguousarray(l[0], np.uint64) for l in layers]
        cis = [np.ascontiguousarray(l[1] if len(l[1]) else np.zeros(1, np.uint32), np.uint32) for l in layers]
        RP = (C.c_void_p * nl)(*[a.ctypes.data for a in rps])
        CI = (C.c_void_p * nl)(*[a.ctypes.data for a in cis])
        self.h = C.c_void_p()
        L.check(L.lib().sdb_hnsw_load(ctx.h, self.dim, L.METRIC[self.metric], self.n, C.c_void_p(vec.ctypes.data), nl,
                                      RP, CI, int(entry_point), C.byref(self.h)))

    @classmethod
    def from_kv(cls, ctx, dim, state_value, he_items, hn_items_per_layer, metric="EUCLIDEAN", elem_docs=None):
        """Loads the index straight from raw KV values (staging.py): `state_value` = the Hs value, `he_items` =
        [(element id, He value)], `hn_items_per_layer[l]` = [(node id, Hn value)] of layer l (0 first), each in key
        order.  Decoding happens on the GPU (sdb_hnsw_load_staged).  Mirrors Hnsw::check_state + HnswLayer::load
        (hnsw/mod.rs:187-224, hnsw/layer.rs:505-560) for indexes without legacy Hl chunks."""
        from . import staging as S
        st = S.parse_hnsw_state(state_value)
        if st["layer0"]["chunks"] or any(l["chunks"] for l in st["layers"]):
            raise L.SdbError(L.SDB_EUNSUPPORTED, "legacy Hl chunks present: run the reference's migration first")
        nl = 1 + len(st["layers"])
        if len(hn_items_per_layer) != nl:
            raise L.SdbError(L.SDB_EINVAL, f"state names {nl} layers, {len(hn_items_per_layer)} given")
        self = cls.__new__(cls)
        self.ctx, self.metric, self.dim, self.elem_docs = ctx, metric.upper(), int(dim), elem_docs
        self.n = int(st["next_element_id"])
        vb, vo, vi = S.pack_values(he_items)
        packs = [S.pack_values(it) for it in hn_items_per_layer]
        NB = (C.c_void_p * nl)(*[p[0].ctypes.data for p in packs])
        NO = (C.c_void_p * nl)(*[p[1].ctypes.data for p in packs])
        NI = (C.c_void_p * nl)(*[p[2].ctypes.data for p in packs])
        NN = (C.c_uint64 * nl)(*[len(it) for it in hn_items_per_layer])
        self.h = C.c_void_p()
        bad = C.c_uint64(0)
        ep = -1 if st["enter_point"] is None else int(st["enter_point"])
        L.check(L.lib().sdb_hnsw_load_staged(ctx.h, self.dim, L.METRIC[self.metric], self.n, C.c_void_p(vb.ctypes.data),
                                             C.c_void_p(vo.ctypes.data), C.c_void_p(vi.ctypes.data), len(he_items), nl,
                                             NB, NO, NI, NN, ep, C.byref(self.h), C.byref(bad)))
        self.n_bad = bad.value
        return self

    def search_graph(self, queries, k, ef, counters=False, truthy=None):
        """truthy: optional predicate mask, one byte per element (Hnsw::knn_search_with_filter, hnsw/mod.rs:488-515)"""
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:  # Error::InvalidVectorDimension  idx/trees/vector.rs:643-652
            raise L.SdbError(L.SDB_EDIM, f"Incorrect vector dimension ({q.shape[1]}). Expected a vector of {self.dim} dimension.")
        nq = q.shape[0]
        ids = np.zeros((nq, max(k, 1)), np.uint64)
        dist = np.zeros((nq, max(k, 1)), np.float64)
        cnt = np.zeros(nq, np.uint32)
        ctr = np.zeros((nq, 2), np.uint64)
        if truthy is not None:
            t = np.ascontiguousarray(truthy, np.uint8)
            if t.shape != (self.n,):
                raise L.SdbError(L.SDB_EINVAL, f"predicate mask must have one byte per element ({self.n})")
            L.check(L.lib().sdb_hnsw_search_filtered(self.h, C.c_void_p(q.ctypes.data), nq, int(k), int(ef),
                                                     C.c_void_p(t.ctypes.data), C.c_void_p(ids.ctypes.data),
                                                     C.c_void_p(dist.ctypes.data), C.c_void_p(cnt.ctypes.data),
                                                     C.c_void_p(ctr.ctypes.data)))
        else:
            L.check(L.lib().sdb_hnsw_search(self.h, C.c_void_p(q.ctypes.data), nq, int(k), int(ef),
                                            C.c_void_p(ids.ctypes.data), C.c_void_p(dist.ctypes.data),
                                            C.c_void_p(cnt.ctypes.data), C.c_void_p(ctr.ctypes.data)))
        if counters:
            return ids, dist, cnt, ctr
        return ids, dist, cnt

    def knn_search(self, query, k, ef, truthy_docs=None):
        """-> [(doc_id, distance)] ordered by (distance, doc id), at most k  (one query).
        truthy_docs: optional set of doc ids passing the WHERE condition (cond_filter of HnswIndex::knn_search,
        hnsw/index.rs:270-335): an element enters the result if ANY of its docs is truthy
        (HnswTruthyDocumentFilter::check_any_doc_truthy, hnsw/filter.rs:52-62); add_graph_results then adds ALL docs
        of that element (hnsw/index.rs:454-475) -- the executor re-applies the WHERE clause downstream."""
        truthy = None
        if truthy_docs is not None:
            truthy = np.zeros(self.n, np.uint8)
            for e in range(self.n):
                docs = [e] if self.elem_docs is None else self.elem_docs[e]
                truthy[e] = any(d in truthy_docs for d in docs)
        ids, dist, cnt = self.search_graph(np.asarray(query, np.float32)[None, :], k, ef, truthy=truthy)
        res = []  # KnnResultBuilder: BTreeSet<(dist, doc)>, pop the largest when over k; check_add uses `>`
        for j in range(int(cnt[0])):
            d, e = float(dist[0, j]), int(ids[0, j])
            if len(res) >= k and d > res[-1][0]:
                continue
            docs = [e] if self.elem_docs is None else self.elem_docs[e]
            for doc in docs:
                res.append((d, int(doc)))
                res.sort()
                if len(res) > k:
                    res.pop()
        return [(doc, d) for d, doc in res]

    def close(self):
        if self.h:
            L.lib().sdb_hnsw_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
