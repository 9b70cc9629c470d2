"""dev: recall of the GPU batch-built graph vs the reference-style (oracle, serial insertion) graph, same data."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import pyoracle as O
from surrealdb_b200 import Context, HnswIndex
from surrealdb_b200.hnsw_build import build_layers
ctx = Context(0)
n, dim, nq = 20000, 32, 500
g = torch.Generator(device="cuda").manual_seed(3)
for name, sigma, ncl in (("tight", 0.03, 256), ("clustered", 0.3, 64)):
    if ncl:
        centers = torch.randn((ncl, dim), generator=g, device="cuda")
        x = (centers[torch.randint(0, ncl, (n,), generator=g, device="cuda")] + sigma * torch.randn((n, dim), generator=g, device="cuda")).contiguous()
        q = (centers[torch.randint(0, ncl, (nq,), generator=g, device="cuda")] + sigma * torch.randn((nq, dim), generator=g, device="cuda"))
    else:
        x = torch.rand((n, dim), generator=g, device="cuda") * 2 - 1
        q = torch.rand((nq, dim), generator=g, device="cuda") * 2 - 1
    xh, qh = x.cpu().numpy(), q.cpu().numpy()
    truth = [set(O.vec_knn_f32(xh, qh[i], "euclidean", 10)[0].tolist()) for i in range(nq)]
    t0 = time.time()
    h = O.Hnsw(dim, "euclidean", m=16, efc=150, seed=1)
    for v in xh:
        h.insert(v)
    gref = h.export()
    t_ref = time.time() - t0
    graphs = {"reference-style (oracle insert)": (gref["layers"], gref["entry_point"])}
    for heur, prefix in ((True, False), (True, True)):
        t0 = time.time()
        layers, entry, _ = build_layers(ctx, x, n, dim, "EUCLIDEAN", m=16, m0=32, seed=5, heuristic=heur, prefix=prefix)
        graphs[f"gpu batch heuristic={heur} prefix={prefix} ({time.time()-t0:.1f}s)"] = (layers, entry)
    for gname, (layers, entry) in graphs.items():
        idx = HnswIndex(ctx, xh, layers, entry, "EUCLIDEAN")
        deg = np.diff(layers[0][0].astype(np.int64)).mean()
        for ef in (16, 64, 200):
            ids, dist, cnt, ctr = idx.search_graph(qh, 10, ef, counters=True)
            rec = np.mean([len(truth[i] & set(ids[i, :cnt[i]].tolist())) / 10 for i in range(nq)])
            print(f"{name:10s} {gname:45s} deg0={deg:5.1f} ef={ef:3d} recall={rec:.3f} visited={ctr[:,0].mean():.0f}", flush=True)
    print(f"(oracle build took {t_ref:.1f}s)")
