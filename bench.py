#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native KNN hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA through the C ABI)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

Workload (config.workload): exact brute-force cosine KNN, k=10, batch of 1024 f64 queries per step over a
10M x 768 f32 corpus (the configuration BASELINE.json's metric is quoted on; it fits one B200).  With N
GPUs the SAME 10M-row corpus is row-sharded N ways (strong scaling), each rank screens + exactly re-ranks
its shard, one NCCL all-gather moves the per-shard top-k, and a merge kernel produces the global top-k.
Data are synthetic: a counter-based generator produces identical values on every GPU and on the CPU.

One JSON line is printed by rank 0 (see README / the driver contract for the keys).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows, dim, batch, k)
    "northstar_10Mx768_b1024_k10_cosine_bruteforce": (10_000_000, 768, 1024, 10),
    "c2_1Mx768_b1024_k10_cosine_bruteforce": (1_000_000, 768, 1024, 10),
    "tiny_100kx128_b64_k10_cosine_bruteforce": (100_000, 128, 64, 10),
    # BASELINE config 4 (quoted there on 8 GPUs; 107 GB of corpus + screen copies still fit one B200)
    "c4_10Mx1536_b4096_k100_cosine_bruteforce": (10_000_000, 1536, 4096, 100),
}
SEED_CORPUS = 0x5DB00002
SEED_QUERY = 0x5DB0A000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                          ("sw_power_cap", 6)):
            if any(len(r) >= 7 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_sample(rows_total, dim, budget_s, threads):
    """bounded sample of the workload for the CPU arm: the first 262144 rows, and as many queries (a multiple
    of the thread count) as fit the time budget"""
    from surrealdb_b200.synthetic import gen_f32
    sample_rows = int(min(rows_total, 262_144))
    nq = None  # decided by cpu_baseline's calibration run
    corpus = np.empty((sample_rows, dim), np.float32)
    step = 1 << 16
    for r0 in range(0, sample_rows, step):
        r1 = min(sample_rows, r0 + step)
        corpus[r0:r1] = gen_f32(SEED_CORPUS, r0 * dim, (r1 - r0) * dim).reshape(r1 - r0, dim)
    queries = gen_f32(SEED_QUERY, 0, 64 * threads * dim).reshape(64 * threads, dim).astype(np.float64)
    return corpus, queries, budget_s


def cpu_baseline(rows_total, dim, k, budget_s=15.0, threads=None, sample=None):
    """The ONE place bench.py touches oracle/ (CPU baseline / reference arm).  Times the oracle -- a faithful
    port of the reference's f64 Vec<Number> distance + KnnTopK selection -- on the host cores, on a bounded
    sample; per-query cost is linear in rows, so the figure is scaled by sample_rows / rows_total."""
    from oracle import pyoracle as O
    threads = threads or os.cpu_count() or 1
    corpus, queries, budget_s = sample if sample is not None else cpu_sample(rows_total, dim, budget_s, threads)
    sample_rows = corpus.shape[0]
    t0 = time.perf_counter()  # calibration: one query per thread tells how many rounds fit the budget
    O.knn_topk_batch(corpus, queries[:threads], "cosine", k, threads)
    t_cal = time.perf_counter() - t0
    rounds = int(max(1, min(64, budget_s // max(t_cal, 1e-3))))
    nq = rounds * threads
    queries = queries[:nq]
    t0 = time.perf_counter()
    O.knn_topk_batch(corpus, queries, "cosine", k, threads)
    dt = time.perf_counter() - t0
    qps = nq / dt * (sample_rows / rows_total)
    return {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{nq} queries x first {sample_rows} of {rows_total} rows in {dt:.2f}s on {threads} threads, "
                      f"scaled by rows ({sample_rows}/{rows_total}); the port omits the reference's KV scan + "
                      "document decode, so it is an optimistic stand-in for the Rust path"}, dt


def run_reference(args, rows, dim, batch, k, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    threads = os.cpu_count() or 1
    sample = cpu_sample(rows, dim, 4.0, threads)
    vals, last = [], None
    t_all0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_baseline(rows, dim, k, threads=threads, sample=sample)
        if i >= args.warmup:
            vals.append(cb["value"])
            last = cb
    qps = float(np.mean(vals))
    last["value"] = qps
    out = {"impl": "reference", "metric": "KNN queries/sec @recall@10=1.0 (exact brute force)", "value": qps,
           "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": batch / qps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": {"workload": wl, "rows": rows, "dim": dim, "batch": batch,
                                                            "k": k, "metric": "cosine"},
           "cpu_baseline": last, "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                                         "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="northstar_10Mx768_b1024_k10_cosine_bruteforce", choices=list(WORKLOADS))
    ap.add_argument("--screen", default="AUTO")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rows, dim, batch, k = WORKLOADS[args.workload]
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, rows, dim, batch, k, args.workload)

    import torch
    import torch.distributed as dist
    from surrealdb_b200 import Context, VectorColumn
    from surrealdb_b200.engine import shard_block_layout, topk_merge_device
    from surrealdb_b200.synthetic import gen_f32

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly ONE JSON line: NCCL_DEBUG=VERSION would print a banner there at communicator creation
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- shard the corpus row-wise (contiguous blocks, tile aligned); global row id = base + local ----
    from surrealdb_b200.sharding import shard_range
    base, n_local = shard_range(rows, world, rank)
    ctx = Context(local)
    col = VectorColumn(ctx, dim, "COSINE", "F32", capacity=max(n_local, 1))
    chunk = 1 << 20
    for r0 in range(0, n_local, chunk):
        col.append_synthetic(SEED_CORPUS, base + r0, min(chunk, n_local - r0))
    col.finalize()
    col.set_screen(args.screen)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    # ---- query batches: generated once on the host (pinned, f64), a device-resident copy for `value` ----
    n_batches = args.warmup + args.steps
    q_host = [torch.from_numpy(gen_f32(SEED_QUERY + b, 0, batch * dim).reshape(batch, dim).astype(np.float64)).pin_memory()
              for b in range(n_batches)]
    q_dev = [q.to(dev) for q in q_host]
    # one rank's result block = rows u64 | dist f64 | count u32, contiguous, so ONE all-gather moves it
    off_rows, off_dist, off_cnt, blk = shard_block_layout(batch, k)
    block = torch.zeros((blk,), dtype=torch.uint8, device=dev)
    o_rows = block[off_rows:off_dist].view(torch.int64).view(batch, k)
    o_dist = block[off_dist:off_cnt].view(torch.float64).view(batch, k)
    o_cnt = block[off_cnt:off_cnt + batch * 4].view(torch.int32)
    if world > 1:
        gathered = torch.zeros((world * blk,), dtype=torch.uint8, device=dev)
        f_rows, f_dist, f_cnt = torch.zeros_like(o_rows), torch.zeros_like(o_dist), torch.zeros_like(o_cnt)
    h_rows = torch.zeros((batch, k), dtype=torch.int64).pin_memory()
    h_dist = torch.zeros((batch, k), dtype=torch.float64).pin_memory()

    def step_device(qd):
        col.knn_device(qd.data_ptr(), batch, k, base, o_rows.data_ptr(), o_dist.data_ptr(), o_cnt.data_ptr())
        if world > 1:  # ONE all-gather of the per-shard top-k blocks, then the merge kernel on every rank
            dist.all_gather_into_tensor(gathered, block)
            torch.cuda.current_stream().synchronize()
            gp = gathered.data_ptr()
            topk_merge_device(ctx, world, batch, k, gp + off_rows, gp + off_dist, gp + off_cnt,
                              f_rows.data_ptr(), f_dist.data_ptr(), f_cnt.data_ptr(),
                              stride_rows=blk // 8, stride_dist=blk // 8, stride_counts=blk // 4)
            return f_rows, f_dist
        return o_rows, o_dist

    def step_e2e(qh, d_q):
        d_q.copy_(qh, non_blocking=True)  # H2D of this step's queries (pinned)
        torch.cuda.current_stream().synchronize()
        r, d = step_device(d_q)
        h_rows.copy_(r, non_blocking=True)  # D2H of the result
        h_dist.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, label):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        sync_all()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    # ---- warm-up ----
    for b in range(args.warmup):
        step_device(q_dev[b])
    # ---- timed: device-resident inputs (`value`) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    screen_ms, total_ms, fallbacks = [], [], 0

    def run_value():
        nonlocal fallbacks
        for b in range(args.warmup, n_batches):
            step_device(q_dev[b])
            s = col.stats()
            screen_ms.append(s["screen_ms"])
            total_ms.append(s["total_ms"])
            fallbacks += s["n_fallback"]

    ms_value, wall_value = timed(run_value, "value")
    launches = ctx.kernel_launches() - launches0
    # ---- timed: end to end through the host-buffer path (`e2e`) ----
    d_q = torch.empty_like(q_dev[0])
    step_e2e(q_host[0], d_q)

    def run_e2e():
        for b in range(args.warmup, n_batches):
            step_e2e(q_host[b], d_q)

    ms_e2e, wall_e2e = timed(run_e2e, "e2e")
    clocks = sampler.stop() if rank == 0 else None
    stats = col.stats()

    # ---- HBM-bound regime (small batches), reported next to the headline: f32 streaming kernel and bf16 screen ----
    hbm_regime = []
    if world == 1:
        for scr, b in (("SIMT_F32", 1), ("SIMT_F32", 8), ("TC_BF16", 16), ("TC_INT8", 16)):
            col.set_screen(scr)
            qd = q_dev[0][:b].contiguous()
            best = None
            for _ in range(4):
                col.knn_device(qd.data_ptr(), b, k, base, o_rows.data_ptr(), o_dist.data_ptr(), o_cnt.data_ptr())
                s = col.stats()
                best = s if best is None or s["screen_ms"] < best["screen_ms"] else best
            byts = n_local * (dim * {"SIMT_F32": 4.0, "TC_BF16": 2.0, "TC_INT8": 1.0}[scr] + 4.0) + b * dim * 4.0
            hbm_regime.append({"screen": scr, "batch": b, "screen_ms": best["screen_ms"], "total_ms": best["total_ms"],
                               "algorithmic_bytes": byts, "GBps": byts / (best["screen_ms"] * 1e-3) / 1e9,
                               "qps": b / (best["total_ms"] * 1e-3)})
        col.set_screen(args.screen)

    if rank == 0:
        pk = peaks()
        qps = batch * args.steps / (ms_value * 1e-3)
        qps_e2e = batch * args.steps / (ms_e2e * 1e-3)
        scr_ms = float(np.mean(screen_ms))
        n_shard = n_local
        screen_name = {1: "SIMT_F32", 2: "TC_BF16", 3: "NONE_EXACT", 4: "TC_INT8"}.get(stats["screen_used"], "?")
        if stats["screen_used"] in (2, 4):
            flops = 2.0 * batch * n_shard * dim
            ach = flops / (scr_ms * 1e-3) / 1e12
            i8 = stats["screen_used"] == 4
            peak = pk["bf16_tflops_sustained"] * (2.0 if i8 else 1.0)
            roof = {"bound": "tensor", "kernel": "screen_tc_kernel<cosine,int8> (tcgen05 kind::i8)" if i8 else "screen_tc_kernel (tcgen05 kind::f16 bf16)",
                    "achieved": ach, "peak": peak, "unit": "TOP/s" if i8 else "TFLOP/s", "frac": ach / peak,
                    "peak_source": pk["source"] + (" (2 x sustained cuBLAS bf16: the int8 tensor rate is twice the bf16 rate on B200; "
                                                   "MEASURED_PEAKS.json has no int8 figure)" if i8 else " (sustained cuBLAS bf16)"),
                    "traffic": None, "algorithmic_flops_per_launch": flops,
                    "launch_note": f"the screen runs as {stats['n_passes']} launches of this kernel per step (threshold-refinement passes "
                                   "over disjoint tile subsets); 'achieved' = flops of all of them / CUDA-event time of the whole screen "
                                   "phase on the library stream (includes the compaction kernels between passes)"}
            try:  # DRAM traffic of the dominant launch, from the committed ncu capture (not re-measured here)
                tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["screen_tc_int8" if i8 else "screen_tc_bf16"]
                if rows == 10_000_000 and world == 1:
                    roof["traffic"] = tr["bytes"]
                    roof["traffic_note"] = ("dram read+write bytes of ONE launch, ncu --set full: " + tr.get("launch", "largest pass launch")
                                            + "; " + tr["source"] + f"; algorithmic bytes of that launch {tr['algorithmic_bytes_same_launch']:.4g}")
            except Exception:
                pass
        else:
            passes_over_corpus = (batch + 7) // 8
            byts = passes_over_corpus * (n_shard * dim * 4.0 + n_shard * 4.0) + batch * dim * 4.0
            ach = byts / (scr_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "screen_simt_kernel (f32 stream, 8 queries per corpus pass)",
                    "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "peak_source": pk["source"] + " (copy bandwidth)", "traffic": None,
                    "algorithmic_bytes_per_step": byts}
        out = {"metric": "KNN queries/sec @recall@10=1.0 (exact brute force)", "value": qps, "unit": "queries/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_value / args.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": {4: "s8 screen + f64 exact", 2: "bf16 screen + f64 exact", 1: "f32 screen + f64 exact"}.get(stats["screen_used"], "f64"),
               "data": "synthetic",
               "config": {"workload": args.workload, "rows": rows, "dim": dim, "batch": batch, "k": k,
                          "metric": "cosine", "corpus_dtype": "f32 master + bf16 and int8 screen copies",
                          "screen": screen_name, "exact_rerank": "f64 sequential (reference arithmetic)",
                          "sharding": f"rows/{world}", "l2": "corpus shard (>= 3.8 GB) is larger than L2; no flush needed",
                          "fallback_queries_in_timed_region": int(fallbacks)},
               "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": batch * dim * 8,
                       "d2h_bytes_per_step": batch * k * 16, "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
               "timing": {"value_ms_events": ms_value, "value_ms_wall": wall_value, "e2e_ms_events": ms_e2e,
                          "e2e_ms_wall": wall_e2e, "lib_total_ms_mean": float(np.mean(total_ms)),
                          "lib_screen_ms_mean": scr_ms}}
        for h in hbm_regime:
            h["frac_of_measured_hbm_peak"] = h["GBps"] / pk["hbm_gbs"]
        out["hbm_bound_regime"] = hbm_regime
        if world == 1 and not args.no_cpu_baseline:
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
            out["cpu_baseline"], _ = cpu_baseline(rows, dim, k)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
