#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== hnsw tests"; timeout 900 python -m pytest tests/test_gpu_hnsw.py -q --timeout 600 > gpurun_out/c11_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/c11_tests.log
echo "== hnsw 1M x 768 incremental"; timeout 900 python bench_extra.py hnsw --rows 1000000 --dim 768 --metric cosine --builder incremental --efc 150 --queries 10000 > gpurun_out/c11_hnsw_1M.json 2> gpurun_out/c11_hnsw_1M.err; echo "rc=$?"; tail -5 gpurun_out/c11_hnsw_1M.err; cat gpurun_out/c11_hnsw_1M.json
echo "== hnsw 1M x 768 batch"; timeout 900 python bench_extra.py hnsw --rows 1000000 --dim 768 --metric cosine --builder batch --queries 10000 --no-cpu > gpurun_out/c11_hnsw_1M_batch.json 2> gpurun_out/c11_hnsw_1M_batch.err; echo "rc=$?"; tail -3 gpurun_out/c11_hnsw_1M_batch.err; cat gpurun_out/c11_hnsw_1M_batch.json
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; echo "rc=$?"; cat gpurun_out/c11_bench.json
