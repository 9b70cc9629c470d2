#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c9_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/c9_tests.log
echo "== shard perf 1.25M"; timeout 600 python scripts/shard_perf.py 1250000 > gpurun_out/c9_shard_perf.log 2>&1; cat gpurun_out/c9_shard_perf.log
echo "== shard perf 10M"; timeout 600 python scripts/shard_perf.py 10000000 > gpurun_out/c9_shard_perf_10M.log 2>&1; cat gpurun_out/c9_shard_perf_10M.log
echo "== bench clustered"; timeout 900 python bench.py --steps 10 --warmup 3 --data clustered --no-cpu-baseline --no-extras > gpurun_out/c9_bench_clustered.json 2> gpurun_out/c9_bench_clustered.err; echo "rc=$?"; head -c 300 gpurun_out/c9_bench_clustered.json; echo
echo "== bench c4"; timeout 900 python bench.py --steps 5 --warmup 3 --workload c4_10Mx1536_b4096_k100_cosine_bruteforce --no-cpu-baseline --no-extras > gpurun_out/c9_bench_c4.json 2> gpurun_out/c9_bench_c4.err; echo "rc=$?"; head -c 300 gpurun_out/c9_bench_c4.json; echo
