"""Host-side mirror of the reference's operator interface for the hot path.

The reference is Rust (no cargo/rustc in this image), so the operator glue a maintainer would add behind
a `gpu-knn` cargo feature is mirrored here in Python with the same names, argument
meaning and error behaviour, on top of the same C ABI:

  KnnTopK(input, field, query_vector, k, distance)   core/exec/operators/knn_topk.rs:100-118
      .execute()  -> records nearest-first           knn_topk.rs:166-267
      .name() / .attrs()                              knn_topk.rs:133-144   (EXPLAIN output)
  KnnScan(index, vector, k, ef, table_name, ...)      core/exec/operators/scan/knn.rs:68-118,135-347 (HNSW-backed)
  KnnContext: record id -> distance hand-back         core/exec/function/index.rs:289-314

Records are dicts with an "id"; `input` is any iterable yielding them in scan (record-key) order, i.e.
what TableScan yields.
"""
import numpy as np

from .engine import Context, VectorColumn


class Distance:
    """catalog::Distance (catalog/schema/index.rs:247-284); Debug names as printed by EXPLAIN."""
    Cosine = "Cosine"
    Euclidean = "Euclidean"
    Manhattan = "Manhattan"      # the next four are ranked by the exact kernel only (no screen)
    Chebyshev = "Chebyshev"
    Hamming = "Hamming"
    Pearson = "Pearson"
    Minkowski = "Minkowski"      # exact kernel, order via VectorColumn.set_minkowski_order (pow(): ~1e-14 relative)
    Jaccard = "Jaccard"          # exact kernel, set semantics over the values


class KnnContext(dict):
    """rid -> distance of the last KNN operator (exec/function/index.rs:289-314); backs
    vector::distance::knn()."""

    def insert(self, rid, distance):
        self[rid] = distance


def _pick(value, field):
    """Value::pick for a dotted idiom path; missing parts yield None."""
    cur = value
    for part in field.split("."):
        if not isinstance(cur, dict) or part not in cur:
            return None
        cur = cur[part]
    return cur


def extract_vector(value, field):
    """knn_topk.rs:274-288: Some(vec) only for a non-empty array whose elements are all numbers."""
    arr = _pick(value, field)
    if not isinstance(arr, (list, tuple)) or len(arr) == 0:
        return None
    out = []
    for v in arr:
        if isinstance(v, bool) or not isinstance(v, (int, float, np.integer, np.floating)):
            return None
        out.append(float(v))
    return out


class KnnTopK:
    """Brute-force KNN operator backed by the GPU column.  Pipeline-breaking: consumes the whole input,
    returns the k nearest records ordered by (distance, scan position)."""

    def __init__(self, input, field, query_vector, k, distance, ctx=None):
        self.input = input
        self.field = field
        self.query_vector = [float(x) for x in query_vector]
        self.k = int(k)
        self.distance = distance
        self.knn_context = None
        self._ctx = ctx
        self._column = None
        self._records = None

    def with_knn_context(self, knn_context):
        self.knn_context = knn_context
        return self

    def name(self):
        return "KnnTopK"

    def attrs(self):
        return [("field", self.field), ("k", str(self.k)), ("distance", self.distance),
                ("dimension", str(len(self.query_vector)))]

    def _stage(self):
        """TableScan -> pinned rows -> HBM column (the staging a Rust shim caches per table version)."""
        dim = len(self.query_vector)
        recs, rows, skip = [], [], []
        for rec in self.input:
            vec = extract_vector(rec, self.field)
            recs.append(rec)
            if vec is None or len(vec) != dim:  # extract_vector None, or compute() Err on dimension mismatch
                rows.append([0.0] * dim)
                skip.append(1)
            else:
                rows.append(vec)
                skip.append(0)
        self._records = recs
        if not recs:
            return None
        arr = np.asarray(rows, np.float64)
        f32 = arr.astype(np.float32)
        dtype = "F32" if np.array_equal(f32.astype(np.float64), arr, equal_nan=True) else "F64"
        if self._ctx is None:
            self._ctx = Context(0)
        col = VectorColumn(self._ctx, dim, self.distance.upper(), dtype, capacity=len(recs))
        col.append(f32 if dtype == "F32" else arr)
        if any(skip):
            col.set_skip(np.asarray(skip, np.uint8))
        col.finalize()
        return col

    def execute(self):
        if self._column is None:
            self._column = self._stage()
        if self._column is None or self.k == 0:
            return []
        rows, dist, cnt = self._column.knn(np.asarray([self.query_vector], np.float64), self.k)
        out = []
        for j in range(int(cnt[0])):
            rec = self._records[int(rows[0, j])]
            if self.knn_context is not None and isinstance(rec, dict) and "id" in rec:
                self.knn_context.insert(rec["id"], float(dist[0, j]))
            out.append(rec)
        return out


class KnnBruteForceLegacy(KnnTopK):
    """The legacy two-pass brute force of the old executor: QueryExecutor::knn (idx/planner/executor.rs:283-311)
    feeds every row's distance to a KnnPriorityList (idx/planner/knn.rs:11-106); the second pass re-iterates the
    table and keeps the rows the list retained, so the result comes back in TABLE order, with the distances served
    by vector::distance::knn() (fnc/vector.rs:79-101) from KnnBruteForceResults::get_dist.

    The retained set is the k nearest; when several rows tie at the k-th distance the reference keeps an arbitrary
    subset of that tie group (`HashSet` iteration order, knn.rs:85-93).  This mirror resolves the tie by scan order
    (a valid outcome of the reference), which makes it the GPU top-k followed by a re-sort on scan position."""

    def name(self):
        return "KnnBruteForce"

    def execute(self):
        if self._column is None:
            self._column = self._stage()
        if self._column is None or self.k == 0:
            return []
        rows, dist, cnt = self._column.knn(np.asarray([self.query_vector], np.float64), self.k)
        n = int(cnt[0])
        order = np.argsort(rows[0, :n], kind="stable")
        out = []
        for j in order:
            rec = self._records[int(rows[0, j])]
            if self.knn_context is not None and isinstance(rec, dict) and "id" in rec:
                self.knn_context.insert(rec["id"], float(dist[0, j]))
            out.append(rec)
        return out


class KnnScan:
    """HNSW-backed KNN scan operator: `WHERE emb <|k,ef|> $q` with an HNSW index (scan/knn.rs:68-118,135-347).
    `index` is the device-resident index (surrealdb_b200.hnsw.HnswIndex) with `.name`; `records` maps a vector id
    (doc id / record key) to the record the scan yields (HnswDocs::get_thing + fetch_and_filter_records_batch).
    residual_cond: optional predicate over records pushed into the search so that rows failing it do not consume top-k
    slots (scan/knn.rs:265-273 -> HnswIndex::knn_search cond_filter)."""

    def __init__(self, index, vector, k, ef, table_name, records, knn_context=None, residual_cond=None,
                 index_name=None, state_value=None):
        self.index, self.vector = index, [float(x) for x in vector]
        self.k, self.ef, self.table_name = int(k), int(ef), table_name
        self.records, self.knn_context, self.residual_cond = records, knn_context, residual_cond
        self.index_name = index_name or getattr(index, "name", "idx")
        self.state_value = state_value

    def name(self):
        return "KnnScan"

    def attrs(self):  # scan/knn.rs:110-117
        return [("index", self.index_name), ("k", str(self.k)), ("ef", str(self.ef)),
                ("dimension", str(len(self.vector)))]

    def cardinality_hint(self):  # CardinalityHint::Bounded(k)
        return ("Bounded", self.k)

    def access_mode(self):
        return "ReadOnly"

    def execute(self):
        from . import _lib as L
        # check_state: the device copy must match the persisted layer versions (scan/knn.rs:258-262)
        if self.state_value is not None and not self.index.check_state(self.state_value):
            raise L.SdbError(L.SDB_EINVAL, "Failed to check HNSW index state: the device copy is stale (reload it)")
        if len(self.vector) != self.index.dim:  # Vector::check_dimension  idx/trees/vector.rs:643-652
            raise L.SdbError(L.SDB_EDIM, f"Incorrect vector dimension ({len(self.vector)}). Expected a vector of "
                                         f"{self.index.dim} dimension.")
        truthy = None
        if self.residual_cond is not None:
            truthy = {vid for vid, rec in self.records.items() if self.residual_cond(rec)}
        res = self.index.knn_search(self.vector, self.k, self.ef, truthy_docs=truthy)
        out = []
        for vid, dist in res:
            rec = self.records.get(vid)
            if rec is None:  # HnswDocs::get_thing returned None: the doc vanished
                continue
            if self.knn_context is not None:
                self.knn_context.insert(rec["id"], dist)
            out.append(rec)
        return out
