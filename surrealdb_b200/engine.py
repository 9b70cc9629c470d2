"""Thin object wrappers over the C ABI handles (sdb_ctx, sdb_corpus)."""
import ctypes as C

import numpy as np

from . import _lib as L


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class Context:
    """sdb_ctx: one CUDA device."""

    def __init__(self, device=0, _handle=None):
        self.h = C.c_void_p()
        if _handle is not None:
            self.h = C.c_void_p(_handle)
        else:
            L.check(L.lib().sdb_ctx_create(device, C.byref(self.h)))
        self.device = device

    @staticmethod
    def create_multi(devices):
        """one process, several GPUs: contexts sharing one NCCL communicator (sdb_ctx_create_multi)"""
        devs = (C.c_int * len(devices))(*devices)
        out = (C.c_void_p * len(devices))()
        L.check(L.lib().sdb_ctx_create_multi(devs, len(devices), out))
        return [Context(d, _handle=out[i]) for i, d in enumerate(devices)]

    @staticmethod
    def comm_unique_id():
        """128 bytes rank 0 hands to the other ranks (any out-of-band channel) before comm_init_rank"""
        buf = (C.c_uint8 * L.COMM_ID_BYTES)()
        L.check(L.lib().sdb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, nranks, rank, unique_id):
        buf = (C.c_uint8 * L.COMM_ID_BYTES).from_buffer_copy(unique_id)
        L.check(L.lib().sdb_comm_init_rank(self.h, int(nranks), int(rank), buf))

    def comm_size(self):
        return int(L.lib().sdb_comm_size(self.h))

    def stream(self):
        """raw cudaStream_t (int) all kernels of this context run on"""
        return int(L.lib().sdb_ctx_stream(self.h) or 0)

    def cancel(self):
        """raise the context's cancel flag (any thread); running / later calls return SDB_ECANCELLED until cancel_reset"""
        L.lib().sdb_ctx_cancel(self.h)

    def cancel_reset(self):
        L.lib().sdb_ctx_cancel_reset(self.h)

    def kernel_launches(self):
        return int(L.lib().sdb_ctx_kernel_launches(self.h))

    def close(self):
        if self.h:
            L.lib().sdb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VectorColumn:
    """sdb_corpus: device-resident N x D column of one vector field, in scan order."""

    def __init__(self, ctx, dim, metric="COSINE", dtype="F32", capacity=1 << 20):
        self.ctx, self.dim, self.metric, self.dtype = ctx, int(dim), metric.upper(), dtype.upper()
        self.h = C.c_void_p()
        L.check(L.lib().sdb_corpus_create(ctx.h, self.dim, L.DTYPE[self.dtype], L.METRIC[self.metric],
                                          int(capacity), C.byref(self.h)))

    def __len__(self):
        return int(L.lib().sdb_corpus_rows(self.h))

    def append(self, rows):
        npdt = np.float32 if self.dtype == "F32" else np.float64
        rows = np.ascontiguousarray(rows, npdt)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise L.SdbError(L.SDB_EDIM, f"rows must be (n, {self.dim})")
        L.check(L.lib().sdb_corpus_append(self.h, _ptr(rows), rows.shape[0]))

    def append_device(self, dev_ptr, n):
        L.check(L.lib().sdb_corpus_append_device(self.h, C.c_void_p(dev_ptr), int(n)))

    def append_synthetic(self, seed, first_row, n):
        L.check(L.lib().sdb_corpus_append_synthetic(self.h, int(seed), int(first_row), int(n)))

    def read_rows(self, first_row, n):
        """rows [first_row, first_row+n) of the device-resident master copy as a numpy array"""
        out = np.empty((int(n), self.dim), np.float32 if self.dtype == "F32" else np.float64)
        L.check(L.lib().sdb_corpus_read_rows(self.h, int(first_row), int(n), _ptr(out)))
        return out

    def set_skip(self, skip):
        if skip is None:
            L.check(L.lib().sdb_corpus_set_skip(self.h, None, 0))
        else:
            s = np.ascontiguousarray(skip, np.uint8)
            L.check(L.lib().sdb_corpus_set_skip(self.h, _ptr(s), s.size))

    def remove(self, row_ids):
        """tombstone rows (scan positions): excluded from every later search, no re-finalize needed"""
        ids = np.ascontiguousarray(row_ids, np.uint64)
        L.check(L.lib().sdb_corpus_remove(self.h, _ptr(ids), ids.size))

    def finalize(self):
        L.check(L.lib().sdb_corpus_finalize(self.h))

    def set_screen(self, name):
        L.check(L.lib().sdb_corpus_set_screen(self.h, L.SCREEN[name.upper()]))

    def set_minkowski_order(self, order):
        L.check(L.lib().sdb_corpus_set_minkowski_order(self.h, float(order)))

    def set_schedule(self, streaming):
        """True (default): streaming tensor-core screen with in-kernel threshold refinement; False: multi-pass"""
        L.check(L.lib().sdb_corpus_set_schedule(self.h, int(bool(streaming))))

    def set_exact(self, exact):
        """False = opt-in approximate mode (no proof, no exact fallback)"""
        L.check(L.lib().sdb_corpus_set_exact(self.h, int(bool(exact))))

    def knn(self, queries, k, cancel_flag=None):
        """queries (nq, dim) float64 -> (rows u64 (nq,k), dist f64 (nq,k), count u32 (nq,))"""
        q = np.ascontiguousarray(queries, np.float64)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            # Error::InvalidVectorDimension analogue; KnnTopK itself never raises it (it skips rows), the
            # HNSW path does (idx/trees/vector.rs:643-652)
            raise L.SdbError(L.SDB_EDIM, f"query dimension {q.shape[1]} != {self.dim}")
        nq = q.shape[0]
        rows = np.zeros((nq, max(k, 1)), np.uint64)
        dist = np.zeros((nq, max(k, 1)), np.float64)
        cnt = np.zeros(nq, np.uint32)
        cf = None if cancel_flag is None else C.c_void_p(cancel_flag.ctypes.data)
        L.check(L.lib().sdb_knn_bruteforce(self.h, _ptr(q), nq, int(k), _ptr(rows), _ptr(dist), _ptr(cnt), cf))
        return rows[:, :k], dist[:, :k], cnt

    def knn_device(self, d_queries, nq, k, row_base, d_out_rows, d_out_dist, d_out_count):
        """all arguments are raw device pointers (ints); results complete on return."""
        L.check(L.lib().sdb_knn_bruteforce_device(self.h, C.c_void_p(d_queries), int(nq), int(k), int(row_base),
                                                  C.c_void_p(d_out_rows), C.c_void_p(d_out_dist),
                                                  C.c_void_p(d_out_count)))

    # ---- asynchronous batches (device pointers; buffers must stay valid until wait) ----
    def submit_device(self, d_queries, nq, k, row_base, d_out_rows, d_out_dist, d_out_count):
        t = C.c_uint32()
        L.check(L.lib().sdb_knn_submit_device(self.h, C.c_void_p(d_queries), int(nq), int(k), int(row_base),
                                              C.c_void_p(d_out_rows), C.c_void_p(d_out_dist), C.c_void_p(d_out_count),
                                              C.byref(t)))
        return t.value

    def submit_host(self, h_queries, nq, k, h_out_rows, h_out_dist, h_out_count):
        """raw host pointers (ints), pinned for overlap"""
        t = C.c_uint32()
        L.check(L.lib().sdb_knn_submit(self.h, C.c_void_p(h_queries), int(nq), int(k), C.c_void_p(h_out_rows),
                                       C.c_void_p(h_out_dist), C.c_void_p(h_out_count), C.byref(t)))
        return t.value

    def wait(self, ticket):
        L.check(L.lib().sdb_knn_wait(self.h, int(ticket)))

    # ---- row-sharded search (collective over the context's communicator) ----
    def set_row_base(self, row_base):
        L.check(L.lib().sdb_corpus_set_row_base(self.h, int(row_base)))

    def sharded_submit_device(self, d_queries, nq, k, d_out_rows, d_out_dist, d_out_count):
        t = C.c_uint32()
        L.check(L.lib().sdb_knn_sharded_submit_device(self.h, C.c_void_p(d_queries), int(nq), int(k),
                                                      C.c_void_p(d_out_rows), C.c_void_p(d_out_dist),
                                                      C.c_void_p(d_out_count), C.byref(t)))
        return t.value

    def sharded_submit_host(self, h_queries, nq, k, h_out_rows, h_out_dist, h_out_count):
        t = C.c_uint32()
        L.check(L.lib().sdb_knn_sharded_submit(self.h, C.c_void_p(h_queries), int(nq), int(k), C.c_void_p(h_out_rows),
                                               C.c_void_p(h_out_dist), C.c_void_p(h_out_count), C.byref(t)))
        return t.value

    def sharded_wait(self, ticket):
        L.check(L.lib().sdb_knn_sharded_wait(self.h, int(ticket)))

    def project(self, fn, query=None):
        """One value per row of `vector::<fn>(row, query)` in the reference's f64 arithmetic (fnc/vector.rs): fn is a
        metric name ("EUCLIDEAN", "MANHATTAN", ... = vector::distance::*, "COSINE" = 1 - similarity, "PEARSON" =
        vector::similarity::pearson) or "SIMILARITY_COSINE" / "DOT" / "MAGNITUDE"."""
        fn = fn.upper()
        code = L.VECTOR_FN[fn] if fn in L.VECTOR_FN else L.METRIC[fn]
        out = np.zeros(len(self), np.float64)
        q = None
        if query is not None:
            q = np.ascontiguousarray(query, np.float64)
            if q.shape != (self.dim,):  # check_same_dimension  fnc/util/math/vector.rs:23-32
                raise L.SdbError(L.SDB_EDIM, "The two vectors must be of the same dimension.")
        L.check(L.lib().sdb_corpus_project(self.h, _ptr(q) if q is not None else None, code, _ptr(out)))
        return out

    def stats(self):
        s = L.KnnStats()
        L.check(L.lib().sdb_knn_last_stats(self.h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in L.KnnStats._fields_}

    def close(self):
        if self.h:
            L.lib().sdb_corpus_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def topk_merge_device(ctx, n_lists, nq, k, d_rows, d_dist, d_counts, d_out_rows, d_out_dist, d_out_count,
                      stride_rows=0, stride_dist=0, stride_counts=0):
    L.check(L.lib().sdb_topk_merge_device(ctx.h, n_lists, nq, k, C.c_void_p(d_rows), C.c_void_p(d_dist),
                                          C.c_void_p(d_counts), stride_rows, stride_dist, stride_counts,
                                          C.c_void_p(d_out_rows), C.c_void_p(d_out_dist), C.c_void_p(d_out_count)))


def shard_block_layout(nq, k):
    """byte layout of one rank's result block inside the all-gather buffer:
    rows u64[nq*k] | dist f64[nq*k] | count u32[nq] (padded to 16 bytes)."""
    off_rows, off_dist = 0, nq * k * 8
    off_cnt = 2 * nq * k * 8
    size = (off_cnt + nq * 4 + 15) // 16 * 16
    return off_rows, off_dist, off_cnt, size


def knn_sharded_multi(shards, queries, k):
    """one process, N GPUs: shards[i] is the VectorColumn on the i-th context of Context.create_multi"""
    q = np.ascontiguousarray(queries, np.float64)
    nq = q.shape[0]
    rows = np.zeros((nq, max(k, 1)), np.uint64)
    dist = np.zeros((nq, max(k, 1)), np.float64)
    cnt = np.zeros(nq, np.uint32)
    hs = (C.c_void_p * len(shards))(*[s.h for s in shards])
    L.check(L.lib().sdb_knn_sharded_multi(hs, len(shards), _ptr(q), nq, int(k), _ptr(rows), _ptr(dist), _ptr(cnt)))
    return rows[:, :k], dist[:, :k], cnt
