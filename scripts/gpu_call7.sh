#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests (knn/stream)"; timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_knn.py tests/test_gpu_multi.py tests/test_gpu_hnsw.py -q --timeout 600 > gpurun_out/c7_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/c7_tests.log
echo "== shard perf 1.25M"; timeout 600 python scripts/shard_perf.py 1250000 > gpurun_out/c7_shard_perf.log 2>&1; cat gpurun_out/c7_shard_perf.log
echo "== shard perf 10M"; timeout 600 python scripts/shard_perf.py 10000000 > gpurun_out/c7_shard_perf_10M.log 2>&1; cat gpurun_out/c7_shard_perf_10M.log
echo "== launch list stream 1.25M"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c7_launches_stream.csv python scripts/one_shard.py 1250000 stream 5 > gpurun_out/c7_st.log 2>&1; echo rc=$?
