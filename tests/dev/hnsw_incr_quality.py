"""dev (GPU): recall / degree / build time of hnsw_build.build_incremental on the data of hnsw_ref_quality_cpu.py
(same numpy seed => the same rows and queries), for a few builder settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from surrealdb_b200 import Context, HnswIndex
from surrealdb_b200.hnsw_build import build_incremental
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
configs = sys.argv[2:] or ["0.25,0", "0.25,1", "0.1,1"]
dim, nq, sigma = 768, 300, 0.15
ncl = max(1, n // 244)
rng = np.random.default_rng(5)
cent = rng.normal(0, 1, (ncl, dim)); cent /= np.linalg.norm(cent, axis=1, keepdims=True)
x = (cent[rng.integers(0, ncl, n)] + sigma / np.sqrt(dim) * rng.normal(0, 1, (n, dim))).astype(np.float32)
q = (cent[rng.integers(0, ncl, nq)] + sigma / np.sqrt(dim) * rng.normal(0, 1, (nq, dim))).astype(np.float32)
ctx = Context(0)
xd = torch.from_numpy(x).cuda()
qd = torch.nn.functional.normalize(torch.from_numpy(q).cuda().double(), dim=1)
sims = torch.cat([qd @ torch.nn.functional.normalize(xd[i:i + 100000].double(), dim=1).T for i in range(0, n, 100000)], 1)
truth = torch.topk(sims, 10, dim=1).indices.cpu().numpy()
del sims
for cfg in configs:
    growth, settle = cfg.split(",")
    t0 = time.time()
    res = build_incremental(ctx, xd, "COSINE", m=16, m0=32, efc=150, seed=1, growth=float(growth), settle=bool(int(settle)),
                            boot_min=min(65536, max(4096, n // 8)))
    torch.cuda.synchronize()
    bs = time.time() - t0
    idx = HnswIndex.from_device(ctx, res["x"], res["layers_dev"], res["entry"], "COSINE")
    order = res["order"]
    deg = float(res["layers_dev"][0][1].numel()) / n
    for ef in (64, 128):
        ids, dist, cnt, ctr = idx.search_graph(q, 10, ef, counters=True)
        rec = np.mean([len(set(order[ids[i, :cnt[i]].astype(np.int64)].tolist()) & set(truth[i].tolist())) / 10 for i in range(nq)])
        print(f"n={n} growth={growth} settle={settle} build={bs:.1f}s deg0={deg:.1f} ef={ef} recall@10={rec:.4f} visited/q={ctr[:,0].mean():.0f}", flush=True)
    del idx, res
    torch.cuda.empty_cache()
