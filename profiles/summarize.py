"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_bench_r1.csv > profiles/r1_launches_bench.txt
  python profiles/summarize.py full gpurun_out/prof_tc_r1b.ncu-rep > profiles/r1_screen_tc_full.txt
"""
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none   ({path})")
    print(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  total {total/1e6:.3f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v)/total*100:6.2f}%  n={len(v):4d}  total={sum(v)/1e6:10.3f} ms  max={max(v)/1e3:10.1f} us  {k[:90]}")


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "launch__occupancy_limit", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "dram__cycles_active.avg.pct",
        "smsp__average_warp", "launch__shared_mem_per_block", "sm__inst_executed_pipe_fma", "smsp__inst_executed.sum "]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on   ({path})")
    for r in rows[2:]:
        name = r[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print(f"## {name[:100]}")
        for i, a in enumerate(h):
            if any(k in a for k in KEYS):
                print(f"{a:90s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
