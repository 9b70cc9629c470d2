"""ctypes binding of the CPU oracle (oracle/libsdb_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under surrealdb_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsdb_oracle.so")

METRICS = {"chebyshev": 0, "cosine": 1, "euclidean": 2, "hamming": 3, "jaccard": 4, "manhattan": 5,
           "minkowski": 6, "pearson": 7}


def build(force=False):
    src = os.path.join(_HERE, "sdb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Num(C.Structure):
    class _V(C.Union):
        _fields_ = [("f", C.c_double), ("i", C.c_int64)]

    _fields_ = [("tag", C.c_int64), ("v", _V)]

    def py(self):
        return int(self.v.i) if self.tag else float(self.v.f)


def _nums(seq):
    arr = (Num * len(seq))()
    for j, x in enumerate(seq):
        if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
            arr[j].tag = 1
            arr[j].v.i = int(x)
        else:
            arr[j].tag = 0
            arr[j].v.f = float(x)
    return arr


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64p, f64p, f32p, u32p, u8p = (C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_float),
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint8))
        L.orc_f64_cosine_distance.restype = C.c_double
        L.orc_f64_euclidean.restype = C.c_double
        L.orc_f32row_cosine_distance.restype = C.c_double
        L.orc_f32row_euclidean.restype = C.c_double
        L.orc_f64_magnitude.restype = C.c_double
        L.orc_f64_metric.restype = C.c_double
        L.orc_f32row_magnitude.restype = C.c_double
        L.orc_knn_topk.restype = C.c_size_t
        L.orc_nd_dot_f32.restype = C.c_float
        L.orc_nd_sumsq_f32.restype = C.c_float
        for n in ("orc_vec_cosine_f32", "orc_vec_l2_f32", "orc_vec_cosine_f64", "orc_vec_l2_f64", "orc_vec_distance_f32"):
            getattr(L, n).restype = C.c_double
        L.orc_dpq_new.restype = C.c_void_p
        L.orc_dpq_len.restype = C.c_size_t
        L.orc_krb_new.restype = C.c_void_p
        L.orc_krb_collect.restype = C.c_size_t
        L.orc_hnsw_new.restype = C.c_void_p
        L.orc_hnsw_insert.restype = C.c_uint64
        L.orc_hnsw_insert_level.restype = C.c_uint64
        L.orc_hnsw_len.restype = C.c_size_t
        L.orc_hnsw_search.restype = C.c_size_t
        L.orc_hnsw_n_layers.restype = C.c_size_t
        L.orc_hnsw_entry_point.restype = C.c_int64
        L.orc_hnsw_layer_edges.restype = C.c_size_t
        L.orc_hnsw_vectors.restype = C.c_void_p
        L.orc_hnsw_search_csr.restype = C.c_size_t
        L.orc_hnsw_search_csr_filtered.restype = C.c_size_t
        L.orc_vec_knn_f32.restype = C.c_size_t
        L.orc_graph_hop.restype = C.c_uint64
        L.orc_graph_collect.restype = C.c_uint64
        L.orc_gen_f32.restype = C.c_float
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ---------------------------------------------------------------- Number-domain metrics
def num_metric(name, a, b, p=3.0):
    """Distance::compute / vector::* on Vec<Number>; Python ints are Number::Int, floats Number::Float.
    Returns (status, value)."""
    L = lib()
    out = Num()
    fn = {"dot": L.orc_num_dot, "cosine_similarity": L.orc_num_cosine_similarity,
          "cosine_distance": L.orc_num_cosine_distance, "euclidean": L.orc_num_euclidean,
          "manhattan": L.orc_num_manhattan, "chebyshev": L.orc_num_chebyshev, "hamming": L.orc_num_hamming,
          "pearson": L.orc_num_pearson, "jaccard": L.orc_num_jaccard}.get(name)
    A, B = _nums(a), _nums(b)
    if name == "minkowski":
        st = L.orc_num_minkowski(A, C.c_size_t(len(a)), B, C.c_size_t(len(b)), C.c_double(p), C.byref(out))
    else:
        st = fn(A, C.c_size_t(len(a)), B, C.c_size_t(len(b)), C.byref(out))
    return st, (out.py() if st == 0 else None)


def num_magnitude(a):
    out = Num()
    lib().orc_num_magnitude(_nums(a), C.c_size_t(len(a)), C.byref(out))
    return out.py()


def num_cmp(a, b):
    A, B = _nums([a]), _nums([b])
    return lib().orc_num_cmp(A, B)


def f64_cosine_distance(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return lib().orc_f64_cosine_distance(_p(a, C.c_double), _p(b, C.c_double), C.c_size_t(a.size))


def f64_metric(metric, a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return lib().orc_f64_metric(C.c_int(METRICS[metric]), _p(a, C.c_double), _p(b, C.c_double), C.c_size_t(a.size))


def f64_euclidean(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return lib().orc_f64_euclidean(_p(a, C.c_double), _p(b, C.c_double), C.c_size_t(a.size))


# ---------------------------------------------------------------- KnnTopK
def knn_topk(corpus, query, metric, k, skip=None):
    """corpus: (N,D) float32 or float64 C-contiguous; query (D,) float64.  -> (rows u64, dist f64)"""
    corpus = np.ascontiguousarray(corpus)
    assert corpus.dtype in (np.float32, np.float64)
    q = np.ascontiguousarray(query, np.float64)
    n, d = corpus.shape
    rows = np.zeros(max(k, 1), np.uint64)
    dist = np.zeros(max(k, 1), np.float64)
    sk = None
    if skip is not None:
        sk = np.ascontiguousarray(skip, np.uint8)
    c = lib().orc_knn_topk(C.c_void_p(corpus.ctypes.data), C.c_int(corpus.dtype == np.float64), C.c_size_t(n),
                           C.c_size_t(d), _p(sk, C.c_uint8) if sk is not None else None, _p(q, C.c_double),
                           C.c_int(METRICS[metric]), C.c_size_t(k), _p(rows, C.c_uint64), _p(dist, C.c_double))
    return rows[:c].copy(), dist[:c].copy()


def knn_topk_batch(corpus, queries, metric, k, n_threads):
    corpus = np.ascontiguousarray(corpus)
    q = np.ascontiguousarray(queries, np.float64)
    n, d = corpus.shape
    nq = q.shape[0]
    rows = np.zeros((nq, k), np.uint64)
    dist = np.zeros((nq, k), np.float64)
    lib().orc_knn_topk_batch(C.c_void_p(corpus.ctypes.data), C.c_int(corpus.dtype == np.float64), C.c_size_t(n),
                             C.c_size_t(d), _p(q, C.c_double), C.c_size_t(nq), C.c_int(METRICS[metric]),
                             C.c_size_t(k), _p(rows, C.c_uint64), _p(dist, C.c_double), C.c_int(n_threads))
    return rows, dist


# ---------------------------------------------------------------- typed f32 metrics
def vec_distance_f32(metric, a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return lib().orc_vec_distance_f32(C.c_int(METRICS[metric]), _p(a, C.c_float), _p(b, C.c_float), C.c_size_t(a.size))


def vec_distance_f64(metric, a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    fn = lib().orc_vec_cosine_f64 if metric == "cosine" else lib().orc_vec_l2_f64
    return fn(_p(a, C.c_double), _p(b, C.c_double), C.c_size_t(a.size))


def vec_knn_f32(corpus, q, metric, k):
    corpus = np.ascontiguousarray(corpus, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    ids = np.zeros(k, np.uint64)
    dist = np.zeros(k, np.float64)
    c = lib().orc_vec_knn_f32(_p(corpus, C.c_float), C.c_size_t(corpus.shape[0]), C.c_size_t(corpus.shape[1]),
                              C.c_int(METRICS[metric]), _p(q, C.c_float), C.c_size_t(k), _p(ids, C.c_uint64),
                              _p(dist, C.c_double))
    return ids[:c].copy(), dist[:c].copy()


# ---------------------------------------------------------------- queues
class DoublePriorityQueue:
    def __init__(self):
        self.h = C.c_void_p(lib().orc_dpq_new())

    def __del__(self):
        lib().orc_dpq_free(self.h)

    def push(self, d, i):
        lib().orc_dpq_push(self.h, C.c_double(d), C.c_uint64(i))

    def __len__(self):
        return lib().orc_dpq_len(self.h)

    def _pop(self, fn):
        d, i = C.c_double(), C.c_uint64()
        return (d.value, i.value) if fn(self.h, C.byref(d), C.byref(i)) else None

    def pop_first(self):
        return self._pop(lib().orc_dpq_pop_first)

    def pop_last(self):
        return self._pop(lib().orc_dpq_pop_last)

    def peek_first(self):
        return self._pop(lib().orc_dpq_peek_first)

    def peek_last_dist(self):
        d = C.c_double()
        return d.value if lib().orc_dpq_peek_last_dist(self.h, C.byref(d)) else None


class KnnResultBuilder:
    def __init__(self, knn):
        self.knn = knn
        self.h = C.c_void_p(lib().orc_krb_new(C.c_size_t(knn)))

    def __del__(self):
        lib().orc_krb_free(self.h)

    def check_add(self, d):
        return bool(lib().orc_krb_check_add(self.h, C.c_double(d)))

    def add_graph_result(self, d, docs):
        a = np.asarray(docs, np.uint64)
        lib().orc_krb_add(self.h, C.c_double(d), _p(a, C.c_uint64), C.c_size_t(a.size))

    def collect(self):
        dist = np.zeros(self.knn + 1, np.float64)
        doc = np.zeros(self.knn + 1, np.uint64)
        c = lib().orc_krb_collect(self.h, _p(dist, C.c_double), _p(doc, C.c_uint64))
        return [(float(dist[i]), int(doc[i])) for i in range(c)]


# ---------------------------------------------------------------- HNSW
class Hnsw:
    """Hnsw<L0,L> restated (F32).  m0 defaults to 2*m, ml to 1/ln(m) like the DDL defaults
    (syn/parser/stmt/define.rs:1102-1183)."""

    def __init__(self, dim, metric="euclidean", m=12, m0=None, efc=150, ml=None, extend_candidates=False,
                 keep_pruned_connections=False, seed=1):
        import math
        m0 = 2 * m if m0 is None else m0
        ml = 1.0 / math.log(m) if ml is None else ml
        self.dim, self.metric = dim, metric
        self.h = C.c_void_p(lib().orc_hnsw_new(C.c_size_t(dim), C.c_int(METRICS[metric]), C.c_size_t(m),
                                                C.c_size_t(m0), C.c_size_t(efc), C.c_double(ml),
                                                C.c_int(int(extend_candidates) | (int(keep_pruned_connections) << 1)),
                                                C.c_uint64(seed)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_hnsw_free(self.h)
        except Exception:  # interpreter shutdown
            pass

    def insert(self, v, level=None):
        v = np.ascontiguousarray(v, np.float32)
        if level is None:
            return lib().orc_hnsw_insert(self.h, _p(v, C.c_float))
        return lib().orc_hnsw_insert_level(self.h, _p(v, C.c_float), C.c_size_t(level))

    def __len__(self):
        return lib().orc_hnsw_len(self.h)

    def search(self, q, k, ef):
        q = np.ascontiguousarray(q, np.float32)
        ids = np.zeros(max(k, 1), np.uint64)
        dist = np.zeros(max(k, 1), np.float64)
        c = lib().orc_hnsw_search(self.h, _p(q, C.c_float), C.c_size_t(k), C.c_size_t(ef), _p(ids, C.c_uint64),
                                  _p(dist, C.c_double))
        return ids[:c].copy(), dist[:c].copy()

    def counters(self):
        v, e = C.c_uint64(), C.c_uint64()
        lib().orc_hnsw_last_counters(self.h, C.byref(v), C.byref(e))
        return v.value, e.value

    def check_props(self):
        return bool(lib().orc_hnsw_check_props(self.h))

    def export(self):
        """-> dict(vectors (n,dim) f32, entry_point, layers=[(row_ptr u64[n+1], col_idx u32[e]), ...])"""
        n = len(self)
        nl = lib().orc_hnsw_n_layers(self.h)
        vec = np.ctypeslib.as_array(C.cast(lib().orc_hnsw_vectors(self.h), C.POINTER(C.c_float)),
                                    shape=(n, self.dim)).copy()
        layers = []
        for l in range(nl):
            e = lib().orc_hnsw_layer_edges(self.h, C.c_size_t(l))
            rp = np.zeros(n + 1, np.uint64)
            ci = np.zeros(max(e, 1), np.uint32)
            lib().orc_hnsw_export_layer(self.h, C.c_size_t(l), _p(rp, C.c_uint64), _p(ci, C.c_uint32), None)
            layers.append((rp, ci[:e].copy()))
        return {"vectors": vec, "entry_point": lib().orc_hnsw_entry_point(self.h), "layers": layers,
                "metric": self.metric}


def hnsw_search_csr(graph, q, k, ef, truthy=None, all_docs_pending=None):
    """Hnsw::knn_search (truthy=None) / knn_search_with_filter (truthy = one byte per element) over an
    exported/imported graph. -> (ids, dist, (visited, expanded)).  all_docs_pending: the pending-docs bitmap of
    HnswIndex::knn_search evaluated per element (unfiltered search only; the filtered search folds it into truthy)."""
    vec = np.ascontiguousarray(graph["vectors"], np.float32)
    n, dim = vec.shape
    nl = len(graph["layers"])
    rps = [np.ascontiguousarray(l[0], np.uint64) for l in graph["layers"]]
    cis = [np.ascontiguousarray(l[1] if l[1].size else np.zeros(1, np.uint32), np.uint32) for l in graph["layers"]]
    RP = (C.POINTER(C.c_uint64) * nl)(*[_p(a, C.c_uint64) for a in rps])
    CI = (C.POINTER(C.c_uint32) * nl)(*[_p(a, C.c_uint32) for a in cis])
    q = np.ascontiguousarray(q, np.float32)
    ids = np.zeros(max(k, 1), np.uint64)
    dist = np.zeros(max(k, 1), np.float64)
    cnt = np.zeros(2, np.uint64)
    if all_docs_pending is not None and truthy is None:
        t = np.ascontiguousarray(all_docs_pending, np.uint8)
        assert t.size == n
        lib().orc_hnsw_search_csr_pending.restype = C.c_size_t
        c = lib().orc_hnsw_search_csr_pending(_p(vec, C.c_float), C.c_size_t(n), C.c_size_t(dim),
                                              C.c_int(METRICS[graph["metric"]]), C.c_size_t(nl), RP, CI,
                                              C.c_int64(graph["entry_point"]), _p(q, C.c_float), C.c_size_t(k),
                                              C.c_size_t(ef), _p(t, C.c_uint8), _p(ids, C.c_uint64),
                                              _p(dist, C.c_double), _p(cnt, C.c_uint64))
        return ids[:c].copy(), dist[:c].copy(), (int(cnt[0]), int(cnt[1]))
    if truthy is not None:
        t = np.ascontiguousarray(truthy, np.uint8)
        assert t.size == n
        c = lib().orc_hnsw_search_csr_filtered(_p(vec, C.c_float), C.c_size_t(n), C.c_size_t(dim),
                                               C.c_int(METRICS[graph["metric"]]), C.c_size_t(nl), RP, CI,
                                               C.c_int64(graph["entry_point"]), _p(q, C.c_float), C.c_size_t(k),
                                               C.c_size_t(ef), _p(t, C.c_uint8), _p(ids, C.c_uint64),
                                               _p(dist, C.c_double), _p(cnt, C.c_uint64))
        return ids[:c].copy(), dist[:c].copy(), (int(cnt[0]), int(cnt[1]))
    c = lib().orc_hnsw_search_csr(_p(vec, C.c_float), C.c_size_t(n), C.c_size_t(dim), C.c_int(METRICS[graph["metric"]]),
                                  C.c_size_t(nl), RP, CI, C.c_int64(graph["entry_point"]), _p(q, C.c_float),
                                  C.c_size_t(k), C.c_size_t(ef), _p(ids, C.c_uint64), _p(dist, C.c_double),
                                  _p(cnt, C.c_uint64))
    return ids[:c].copy(), dist[:c].copy(), (int(cnt[0]), int(cnt[1]))


# ---------------------------------------------------------------- graph
def graph_hop(row_ptr, col_idx, frontier, limit=0):
    rp = np.ascontiguousarray(row_ptr, np.uint64)
    ci = np.ascontiguousarray(col_idx, np.uint32)
    fr = np.ascontiguousarray(frontier, np.uint32)
    n = lib().orc_graph_hop(_p(rp, C.c_uint64), _p(ci, C.c_uint32), _p(fr, C.c_uint32), C.c_uint64(fr.size),
                            C.c_uint32(limit), None)
    out = np.zeros(max(int(n), 1), np.uint32)
    lib().orc_graph_hop(_p(rp, C.c_uint64), _p(ci, C.c_uint32), _p(fr, C.c_uint32), C.c_uint64(fr.size),
                        C.c_uint32(limit), _p(out, C.c_uint32))
    return out[:n].copy()


def graph_collect(row_ptr, col_idx, start, min_depth=1, max_depth=0, inclusive=False):
    rp = np.ascontiguousarray(row_ptr, np.uint64)
    ci = np.ascontiguousarray(col_idx, np.uint32)
    st = np.ascontiguousarray(start, np.uint32)
    n_nodes = rp.size - 1
    out = np.zeros(n_nodes + st.size + 1, np.uint32)
    n = lib().orc_graph_collect(_p(rp, C.c_uint64), _p(ci, C.c_uint32), C.c_uint64(n_nodes), _p(st, C.c_uint32),
                                C.c_uint64(st.size), C.c_uint32(min_depth), C.c_uint32(max_depth),
                                C.c_int(int(inclusive)), _p(out, C.c_uint32), C.c_uint64(out.size))
    return out[:n].copy()


def graph_recurse_default(row_ptr, col_idx, start, min_depth, max_depth):
    """evaluate_recurse_default (exec/operators/recursion/default.rs:75-133) over `hop`.
    Returns the final frontier array, or None (Value::None)."""
    cur = np.asarray(start, np.uint32)
    depth = 0
    while depth < max_depth:
        nxt = graph_hop(row_ptr, col_idx, cur)
        depth += 1
        if nxt.size == 0 or (nxt.size == cur.size and np.array_equal(nxt, cur)):
            return cur if depth > min_depth else None
        cur = nxt
    return cur if depth >= min_depth else None


# ---------------------------------------------------------------- synthetic data
def gen_f32(seed, first, n):
    out = np.zeros(n, np.float32)
    lib().orc_gen_fill_f32(C.c_uint64(seed), C.c_uint64(first), C.c_uint64(n), _p(out, C.c_float))
    return out


# ---------------------------------------------------------------- legacy KnnPriorityList (idx/planner/knn.rs:11-106)
def knn_priority_list(dists, k):
    """Restatement of the legacy two-pass brute force's first pass (pure Python: small cases only).
    dists: distance per row in scan order (None = row skipped).  Returns (must, tie, left): rows every outcome
    contains, the boundary tie group, and how many of the tie group the reference takes (arbitrary `HashSet` order,
    knn.rs:85-93)."""
    import bisect
    keys, groups, docs = [], {}, set()   # BTreeMap<Number, HashSet<rid>>, docs: HashSet<rid>
    def key(d):
        a = np.float64(d)
        if a == 0:
            a = np.float64(0.0)           # Number::cmp: -0.0 == 0.0
        b = int(a.view(np.uint64))
        return (~b & 0xFFFFFFFFFFFFFFFF) if b >> 63 else (b | (1 << 63))
    for rid, d in enumerate(dists):
        if d is None:
            continue
        kd = key(d)
        if len(docs) >= k and keys and not (keys[-1] > kd):     # check_add  knn.rs:44-52
            continue
        if kd not in groups:                                    # add  knn.rs:54-81
            bisect.insort(keys, kd)
            groups[kd] = [rid]
            docs.add(rid)
        else:
            groups[kd].append(rid)                              # (docs is NOT updated here, as in the reference)
        if len(docs) > k and len(docs) - len(groups[keys[-1]]) >= k:
            for r in groups.pop(keys.pop()):
                docs.discard(r)
    must, left = [], k                                          # build  knn.rs:83-105
    for kd in keys:
        g = groups[kd]
        if len(g) > left:
            return must, g, left
        must += g
        left -= len(g)
        if left == 0:
            break
    return must, [], 0
