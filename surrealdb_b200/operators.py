"""Host-side mirror of the reference's operator interface for the hot path.

The reference is Rust (no cargo/rustc in this image), so the operator glue a maintainer would add behind
a `gpu-knn` cargo feature is mirrored here in Python with the same names, argument
meaning and error behaviour, on top of the same C ABI:

  KnnTopK(input, field, query_vector, k, distance)   core/exec/operators/knn_topk.rs:100-118
      .execute()  -> records nearest-first           knn_topk.rs:166-267
      .name() / .attrs()                              knn_topk.rs:133-144   (EXPLAIN output)
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
    Minkowski = "Minkowski"      # not on the GPU path: VectorColumn raises SDB_EUNSUPPORTED
    Jaccard = "Jaccard"          # idem


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
