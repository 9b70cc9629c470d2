import struct, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from surrealdb_b200 import Context, VectorColumn
rng = np.random.default_rng(6)
corpus = rng.uniform(-1, 1, (500, 24)).astype(np.float32)
corpus[7] = 0.0; corpus[100] = 0.0; corpus[33, 5] = np.nan; corpus[44, 0] = np.inf
queries = rng.uniform(-1, 1, (4, 24)); queries[1] = 0.0; queries[2, 3] = np.nan
ctx = Context(0)
for screen in ("SIMT_F32", "NONE_EXACT"):
    col = VectorColumn(ctx, 24, "COSINE", "F32", capacity=500)
    col.append(corpus); col.finalize(); col.set_screen(screen)
    rows, dist, cnt = col.knn(queries[:1], 499)
    print(screen, col.stats())
    for i in list(range(6)) + list(range(493, 499)):
        print(i, rows[0, i], dist[0, i], hex(struct.unpack('<Q', struct.pack('<d', dist[0, i]))[0]))
