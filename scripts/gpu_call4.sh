#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c4_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/c4_tests.log
echo "== shard perf 1.25M"; timeout 600 python scripts/shard_perf.py 1250000 > gpurun_out/c4_shard_perf.log 2>&1; cat gpurun_out/c4_shard_perf.log
echo "== pipe probe"; timeout 300 python scripts/pipe_probe.py > gpurun_out/c4_pipe_probe.log 2>&1; cat gpurun_out/c4_pipe_probe.log | cut -c1-500
echo "== bench streaming"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench_stream.json 2> gpurun_out/c4_bench_stream.err; echo "rc=$?"; head -c 300 gpurun_out/c4_bench_stream.json; echo
echo "== bench clustered"; timeout 900 python bench.py --steps 10 --warmup 3 --data clustered --no-cpu-baseline --no-extras > gpurun_out/c4_bench_clustered.json 2> gpurun_out/c4_bench_clustered.err; echo "rc=$?"; head -c 300 gpurun_out/c4_bench_clustered.json; echo
echo "== ncu launch list 1.25M-ish"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c4_launches_1M.csv python bench.py --steps 2 --warmup 3 --workload c2_1Mx768_b1024_k10_cosine_bruteforce --no-cpu-baseline --no-extras --no-parity > gpurun_out/c4_ncu_bench.log 2>&1; echo "rc=$?"
