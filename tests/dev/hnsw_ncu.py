"""dev (GPU): one walk launch on a 1M x 768 graph between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
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
idx.search_graph(q[:512], 10, 64)
torch.cuda.synchronize()
torch.cuda.profiler.start()
idx.search_graph(q, 10, 64)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
