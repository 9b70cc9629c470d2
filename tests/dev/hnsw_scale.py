import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from surrealdb_b200 import Context, HnswIndex, VectorColumn
from surrealdb_b200.hnsw_build import build_layers
ctx = Context(0)
dim, nq = 64, 500
g = torch.Generator(device="cuda").manual_seed(3)
for n in (30000, 70000, 150000, 400000):
    x = (torch.rand((n, dim), generator=g, device="cuda") * 2 - 1).contiguous()
    q = torch.rand((nq, dim), generator=g, device="cuda") * 2 - 1
    xh, qh = x.cpu().numpy(), q.cpu().numpy()
    col = VectorColumn(ctx, dim, "EUCLIDEAN", "F32", capacity=n); col.append_device(x.data_ptr(), n); col.finalize()
    rows, _, _ = col.knn(qh.astype(np.float64), 10)
    t0 = time.time()
    layers, entry, levels = build_layers(ctx, x, n, dim, "EUCLIDEAN", m=16, m0=32, seed=5)
    tb = time.time() - t0
    degs = [float(np.diff(l[0].astype(np.int64)).sum()) / max(1, int((levels >= i).sum())) for i, l in enumerate(layers)]
    idx = HnswIndex(ctx, xh, layers, entry, "EUCLIDEAN")
    ids, dist, cnt, ctr = idx.search_graph(qh, 10, 64, counters=True)
    rec = np.mean([len(set(rows[i].tolist()) & set(ids[i, :cnt[i]].tolist())) / 10 for i in range(nq)])
    print(f"n={n} build {tb:.1f}s layers={len(layers)} avgdeg={['%.1f' % d for d in degs]} entry={entry} level={levels[entry]} recall={rec:.3f} visited={ctr[:,0].mean():.0f}", flush=True)
