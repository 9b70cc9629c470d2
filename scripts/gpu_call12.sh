#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== hnsw tests"; timeout 900 python -m pytest tests/test_gpu_hnsw.py -q --timeout 600 > gpurun_out/c12_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/c12_tests.log
echo "== walk perf"; timeout 600 python tests/dev/hnsw_walk_perf.py 1000000 2>&1 | tee gpurun_out/c12_walk_perf.log | tail -8
echo "== quality 60k"; timeout 600 python tests/dev/hnsw_incr_quality.py 60000 0.25,0 0.25,1 0.1,1 2>&1 | tee gpurun_out/c12_quality_60k.log | tail -8
echo "== quality 1M"; timeout 900 python tests/dev/hnsw_incr_quality.py 1000000 0.25,0 0.25,1 0.1,1 2>&1 | tee gpurun_out/c12_quality_1M.log | tail -8
