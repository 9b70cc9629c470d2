"""where does the time go in the pipelined host-buffer path?  (1M x 768, B = 1024)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surrealdb_b200 import Context, VectorColumn
from surrealdb_b200.synthetic import gen_f32
ctx = Context(0)
n, dim, nq, k = 1_000_000, 768, 1024, 10
col = VectorColumn(ctx, dim, "COSINE", "F32", capacity=n)
col.append_synthetic(1, 0, n)
col.finalize()
qs = [torch.from_numpy(gen_f32(100 + b, 0, nq * dim).reshape(nq, dim).astype(np.float64)).pin_memory() for b in range(12)]
outs = [(torch.zeros((nq, k), dtype=torch.int64).pin_memory(), torch.zeros((nq, k), dtype=torch.float64).pin_memory(),
         torch.zeros((nq,), dtype=torch.int32).pin_memory()) for _ in range(2)]
def sub(b, s):
    o = outs[s]
    return col.submit_host(qs[b].data_ptr(), nq, k, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())
for depth in (1, 2, 1, 2):
    torch.cuda.synchronize()
    ts, tw, pend = [], [], []
    t00 = time.perf_counter()
    for b in range(12):
        t0 = time.perf_counter(); pend.append(sub(b, b % 2)); ts.append(time.perf_counter() - t0)
        if len(pend) == depth:
            t0 = time.perf_counter(); col.wait(pend.pop(0)); tw.append(time.perf_counter() - t0)
    while pend:
        t0 = time.perf_counter(); col.wait(pend.pop(0)); tw.append(time.perf_counter() - t0)
    tot = time.perf_counter() - t00
    print(f"depth {depth}: total {tot*1e3:.2f} ms  submit us {[round(x*1e6) for x in ts]}  wait us {[round(x*1e6) for x in tw]}", flush=True)
