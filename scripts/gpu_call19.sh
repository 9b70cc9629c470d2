#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_knn.py tests/test_gpu_stream.py tests/test_gpu_multi.py -q --timeout 600 -x > gpurun_out/c19_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c19_tests.log
SDB_TRACE=1 timeout 300 python scripts/trace_shard.py 1250000 2 14 2> gpurun_out/c19_trace_1250k_d2.txt; echo "rc=$?"
SDB_TRACE=1 timeout 300 python scripts/trace_shard.py 1250000 3 14 2> gpurun_out/c19_trace_1250k_d3.txt; echo "rc=$?"
echo "== shard perf 1.25M"; timeout 600 python scripts/shard_perf.py 1250000 > gpurun_out/c19_shard_perf.log 2>&1; cat gpurun_out/c19_shard_perf.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c19_bench.json
