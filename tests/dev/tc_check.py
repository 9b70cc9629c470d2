"""dev check of the tcgen05 screen against the oracle (run under `timeout` on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import pyoracle as O
from surrealdb_b200 import Context, VectorColumn
from surrealdb_b200.synthetic import gen_f32

ctx = Context(0)
for (n, dim, nq, k) in [(5000, 128, 200, 10), (70000, 768, 300, 10), (3000, 100, 17, 5)]:
    corpus = gen_f32(11, 0, n * dim).reshape(n, dim)
    queries = gen_f32(12, 0, nq * dim).reshape(nq, dim).astype(np.float64)
    for metric in ("COSINE", "EUCLIDEAN"):
        col = VectorColumn(ctx, dim, metric, "F32", capacity=n)
        col.append(corpus); col.finalize(); col.set_screen(os.environ.get("SCREEN", "TC_BF16"))
        t0 = time.time()
        rows, dist, cnt = col.knn(queries, k)
        st = col.stats()
        bad = 0
        for q in range(0, nq, max(1, nq // 25)):
            r, d = O.knn_topk(corpus, queries[q], metric.lower(), k)
            if list(rows[q]) != list(r) or dist[q].tobytes() != d.tobytes():
                bad += 1
                if bad <= 2:
                    print("MISMATCH", q, rows[q], r, dist[q], d)
        print(f"n={n} dim={dim} nq={nq} {metric}: bad={bad} {time.time()-t0:.3f}s stats={st}", flush=True)
