"""dev perf sweep: device-resident brute-force KNN, several batch sizes / screens (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from surrealdb_b200 import Context, VectorColumn
from surrealdb_b200.synthetic import gen_f32

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim, k = 768, 10
ctx = Context(0)
col = VectorColumn(ctx, dim, "COSINE", "F32", capacity=rows)
t0 = time.time()
for r0 in range(0, rows, 1 << 20):
    col.append_synthetic(0x5DB00002, r0, min(1 << 20, rows - r0))
col.finalize()
print(f"staged {rows} rows in {time.time()-t0:.2f}s", flush=True)
dev = torch.device("cuda", 0)
cfgs = (("SIMT_F32", (1, 4, 8)), ("TC_BF16", (16, 128, 1024, 4096)), ("TC_INT8", (16, 128, 1024, 4096)))
if len(sys.argv) > 3:
    cfgs = ((sys.argv[2], (int(sys.argv[3]),)),)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
for screen, batches in cfgs:
    col.set_screen(screen)
    for b in batches:
        q = torch.from_numpy(gen_f32(99, 0, b * dim).reshape(b, dim).astype(np.float64)).to(dev)
        o_r = torch.zeros((b, k), dtype=torch.int64, device=dev)
        o_d = torch.zeros((b, k), dtype=torch.float64, device=dev)
        o_c = torch.zeros((b,), dtype=torch.int32, device=dev)
        ms = []
        if os.environ.get('SDB_TC_DBG'):
            col.set_exact(False)  # ablations produce garbage scores: never run the exact fallback
        elif b > 64:  # guard: a broken screen sends every query to the (slow) exact kernel -- check on a small batch first
            col.knn_device(q.data_ptr(), 64, k, 0, o_r.data_ptr(), o_d.data_ptr(), o_c.data_ptr())
            if col.stats()["n_fallback"] > 6:
                print(f"{screen} B={b}: ABORT, {col.stats()['n_fallback']} of 64 probe queries fell back", flush=True)
                continue
        for it in range(iters):
            col.knn_device(q.data_ptr(), b, k, 0, o_r.data_ptr(), o_d.data_ptr(), o_c.data_ptr())
            ms.append(col.stats())
        s = ms[-1]
        best = min(m["total_ms"] for m in ms[-max(1, len(ms) - 1):])
        bs = min(m["screen_ms"] for m in ms[-max(1, len(ms) - 1):])
        flops = 2.0 * b * rows * dim
        if os.environ.get("SDB_TC_DBG"):
            print("per-call screen_ms:", [round(m["screen_ms"], 3) for m in ms], "passes", [m["n_passes"] for m in ms])
        print(f"{screen} B={b}: total {best:.3f} ms screen {bs:.3f} ms -> {b/best*1e3:.0f} QPS, "
              f"screen {flops/bs/1e9:.1f} TFLOP/s, f32-stream-equiv {rows*dim*4*((b+7)//8 if screen=='SIMT_F32' else 1)/bs/1e6:.0f} GB/s, "
              f"passes {s['n_passes']} fallback {s['n_fallback']} launches {s['kernel_launches']}", flush=True)
