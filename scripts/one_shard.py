"""a few brute-force batches at the shard size, for ncu:  python scripts/one_shard.py rows {stream|multipass} [nbatches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surrealdb_b200 import Context, VectorColumn
from surrealdb_b200.synthetic import gen_f32
rows = int(sys.argv[1]); sched = sys.argv[2]; nb = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dim, nq, k = 768, 1024, 10
ctx = Context(0); dev = torch.device("cuda", 0)
col = VectorColumn(ctx, dim, "COSINE", "F32", capacity=rows)
for r0 in range(0, rows, 1 << 20):
    col.append_synthetic(0x5DB00002, r0, min(1 << 20, rows - r0))
col.finalize(); col.set_schedule(sched == "stream")
o = (torch.zeros((nq, k), dtype=torch.int64, device=dev), torch.zeros((nq, k), dtype=torch.float64, device=dev), torch.zeros((nq,), dtype=torch.int32, device=dev))
for b in range(nb):
    q = torch.from_numpy(gen_f32(0x5DB0A000 + b, 0, nq * dim).reshape(nq, dim).astype(np.float64)).to(dev)
    torch.cuda.synchronize()
    col.knn_device(q.data_ptr(), nq, k, 0, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())
print(col.stats())
