"""dev (GPU): walk-kernel timing on one graph for the occupancy variants (SDB_HNSW_OCC)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from surrealdb_b200 import Context, HnswIndex
from surrealdb_b200.hnsw_build import build_incremental
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dim, nq = 768, 20000
ctx = Context(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0x5DB00003)
centers = torch.nn.functional.normalize(torch.randn((4096, dim), generator=g, device=dev), dim=1)
def sample(cnt):
    c = torch.randint(0, 4096, (cnt,), generator=g, device=dev)
    return (centers[c] + (0.15 / dim ** 0.5) * torch.randn((cnt, dim), generator=g, device=dev)).contiguous()
x = sample(n); q = sample(nq).cpu().numpy()
res = build_incremental(ctx, x, "COSINE", m=16, m0=32, efc=150, seed=7, growth=0.25, settle=False)
idx = HnswIndex.from_device(ctx, res["x"], res["layers_dev"], res["entry"], "COSINE")
for occ in ("6", "8", "4", "6"):
    os.environ["SDB_HNSW_OCC"] = occ
    idx.search_graph(q[:512], 10, 64)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ids, dist, cnt, ctr = idx.search_graph(q, 10, 64, counters=True)
        dt = time.perf_counter() - t0
        best = min(best, dt)
    vis = ctr[:, 0].sum()
    print(f"occ={occ} call {best*1e3:.2f} ms  {nq/best:.0f} QPS  visited/q {vis/nq:.0f}  touched {vis*(4*dim+4)/best/1e12:.2f} TB/s", flush=True)
