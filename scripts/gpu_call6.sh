#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests (knn/stream)"; timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_knn.py tests/test_gpu_multi.py -q --timeout 600 > gpurun_out/c6_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c6_tests.log
echo "== shard perf 1.25M"; timeout 600 python scripts/shard_perf.py 1250000 > gpurun_out/c6_shard_perf.log 2>&1; cat gpurun_out/c6_shard_perf.log
echo "== shard perf 10M"; timeout 600 python scripts/shard_perf.py 10000000 > gpurun_out/c6_shard_perf_10M.log 2>&1; cat gpurun_out/c6_shard_perf_10M.log
echo "== bench clustered"; timeout 900 python bench.py --steps 10 --warmup 3 --data clustered --no-cpu-baseline --no-extras > gpurun_out/c6_bench_clustered.json 2> gpurun_out/c6_bench_clustered.err; echo "rc=$?"; head -c 300 gpurun_out/c6_bench_clustered.json; echo
echo "== launch list stream 1.25M"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c6_launches_stream.csv python scripts/one_shard.py 1250000 stream 5 > gpurun_out/c6_st.log 2>&1; echo rc=$?
