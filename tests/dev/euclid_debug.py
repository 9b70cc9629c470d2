import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from surrealdb_b200 import Context, VectorColumn
from surrealdb_b200.synthetic import gen_f32
ctx = Context(0)
n, dim, nq, k = 5000, 128, 20, 10
corpus = gen_f32(11, 0, n * dim).reshape(n, dim)
queries = gen_f32(12, 0, nq * dim).reshape(nq, dim).astype(np.float64)
for screen in ("SIMT_F32", "TC_BF16"):
    for metric in ("COSINE", "EUCLIDEAN"):
        col = VectorColumn(ctx, dim, metric, "F32", capacity=n)
        col.append(corpus); col.finalize(); col.set_screen(screen)
        rows, dist, cnt = col.knn(queries, k)
        print(screen, metric, col.stats()["n_fallback"], dist[0, :3], flush=True)
