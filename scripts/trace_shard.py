"""SDB_TRACE timeline of pipelined batches at shard size:  SDB_TRACE=1 python scripts/trace_shard.py rows [depth] [nbatches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surrealdb_b200 import Context, VectorColumn
from surrealdb_b200.synthetic import gen_f32
rows = int(sys.argv[1]); depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2; nb = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dim, nq, k = 768, 1024, 10
ctx = Context(0); dev = torch.device("cuda", 0)
col = VectorColumn(ctx, dim, "COSINE", "F32", capacity=rows)
for r0 in range(0, rows, 1 << 20):
    col.append_synthetic(0x5DB00002, r0, min(1 << 20, rows - r0))
col.finalize()
qs = [torch.from_numpy(gen_f32(0x5DB0A000 + b, 0, nq * dim).reshape(nq, dim).astype(np.float64)).to(dev) for b in range(nb)]
outs = [(torch.zeros((nq, k), dtype=torch.int64, device=dev), torch.zeros((nq, k), dtype=torch.float64, device=dev), torch.zeros((nq,), dtype=torch.int32, device=dev)) for _ in range(4)]
torch.cuda.synchronize()
pend = []
for b in range(nb):
    o = outs[b % 4]
    pend.append(col.submit_device(qs[b].data_ptr(), nq, k, 0, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()))
    if len(pend) == depth:
        col.wait(pend.pop(0))
while pend:
    col.wait(pend.pop(0))
