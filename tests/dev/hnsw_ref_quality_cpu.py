"""dev (CPU only): what recall does a REFERENCE-STYLE graph (oracle = restatement of Hnsw::insert, serial insertion) reach
on the C3-shaped data regime (unit-norm centroids + noise of total norm 0.15, ~244 points per cluster, dim 768, cosine,
M=16, efc=150, ef=64, k=10)?  Scaled down to n rows with n/244 clusters so serial insertion finishes in minutes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import pyoracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
dim, nq, sigma = 768, 300, 0.15
ncl = max(1, n // 244)
rng = np.random.default_rng(5)
cent = rng.normal(0, 1, (ncl, dim)); cent /= np.linalg.norm(cent, axis=1, keepdims=True)
x = (cent[rng.integers(0, ncl, n)] + sigma / np.sqrt(dim) * rng.normal(0, 1, (n, dim))).astype(np.float32)
q = (cent[rng.integers(0, ncl, nq)] + sigma / np.sqrt(dim) * rng.normal(0, 1, (nq, dim))).astype(np.float32)
xn = x.astype(np.float64); xn /= np.linalg.norm(xn, axis=1, keepdims=True)
qn = q.astype(np.float64); qn /= np.linalg.norm(qn, axis=1, keepdims=True)
truth = np.argsort(-(qn @ xn.T), axis=1)[:, :10]
t0 = time.time()
h = O.Hnsw(dim, "cosine", m=16, efc=150, seed=1)
for i, v in enumerate(x):
    h.insert(v)
    if i % 10000 == 0:
        print(f"inserted {i} {time.time()-t0:.0f}s", flush=True)
g = h.export()
print(f"build {time.time()-t0:.0f}s layers {len(g['layers'])} deg0 {np.diff(g['layers'][0][0].astype(np.int64)).mean():.1f}", flush=True)
graph = {"vectors": x, "layers": g["layers"], "entry_point": g["entry_point"], "metric": "cosine"}
for ef in (64, 128, 256):
    rec = 0.0
    for i in range(nq):
        ids, _ = O.hnsw_search_csr(graph, q[i], 10, ef)[:2]
        rec += len(set(np.asarray(ids).tolist()) & set(truth[i].tolist())) / 10
    print(f"n={n} clusters={ncl} ef={ef} recall@10={rec/nq:.4f}", flush=True)
