#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_stream.py tests/test_gpu_knn.py tests/test_gpu_multi.py -q --timeout 600 > gpurun_out/c10_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c10_tests.log
echo "== shard perf 1.25M"; timeout 600 python scripts/shard_perf.py 1250000 > gpurun_out/c10_shard_perf.log 2>&1; cat gpurun_out/c10_shard_perf.log
echo "== shard perf 10M"; timeout 600 python scripts/shard_perf.py 10000000 > gpurun_out/c10_shard_perf_10M.log 2>&1; cat gpurun_out/c10_shard_perf_10M.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; echo "rc=$?"; head -c 300 gpurun_out/c10_bench.json; echo
