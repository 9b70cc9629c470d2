#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== hnsw tests"; timeout 900 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_stream.py -q --timeout 600 -x > gpurun_out/c14_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/c14_tests.log
echo "== walk perf"; timeout 600 python tests/dev/hnsw_walk_perf.py 1000000 2>&1 | tee gpurun_out/c14_walk_perf.log | tail -8
echo "== C5 graph N=1"; timeout 900 python bench_extra.py graph --nodes 50000000 --edges 500000000 --checksum --no-cpu > gpurun_out/c14_graph_c5_n1.json 2> gpurun_out/c14_graph_c5_n1.err; echo "rc=$?"; cut -c1-900 gpurun_out/c14_graph_c5_n1.json; tail -3 gpurun_out/c14_graph_c5_n1.err
