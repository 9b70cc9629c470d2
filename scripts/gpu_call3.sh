#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_stream.py tests/test_gpu_knn.py tests/test_gpu_multi.py tests/test_gpu_hnsw.py tests/test_gpu_graph.py -q --timeout 600 > gpurun_out/c3_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/c3_tests.log
echo "== pipe probe"; timeout 300 python scripts/pipe_probe.py > gpurun_out/c3_pipe_probe.log 2>&1; cat gpurun_out/c3_pipe_probe.log | cut -c1-600
echo "== bench streaming"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/c3_bench_stream.json 2> gpurun_out/c3_bench_stream.err; echo "rc=$?"; head -c 300 gpurun_out/c3_bench_stream.json; echo
echo "== bench c2 1M"; timeout 600 python bench.py --steps 20 --warmup 3 --workload c2_1Mx768_b1024_k10_cosine_bruteforce --no-cpu-baseline --no-extras > gpurun_out/c3_bench_c2.json 2> gpurun_out/c3_bench_c2.err; echo "rc=$?"; head -c 300 gpurun_out/c3_bench_c2.json; echo
echo "== bench clustered"; timeout 900 python bench.py --steps 10 --warmup 3 --data clustered --no-cpu-baseline --no-extras > gpurun_out/c3_bench_clustered.json 2> gpurun_out/c3_bench_clustered.err; echo "rc=$?"; head -c 300 gpurun_out/c3_bench_clustered.json; echo
echo "== bench c4 1 gpu"; timeout 900 python bench.py --steps 5 --warmup 3 --workload c4_10Mx1536_b4096_k100_cosine_bruteforce --no-cpu-baseline --no-extras > gpurun_out/c3_bench_c4.json 2> gpurun_out/c3_bench_c4.err; echo "rc=$?"; head -c 300 gpurun_out/c3_bench_c4.json; echo
echo "== ncu launch list 1M"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c3_launches_1M.csv python bench.py --steps 2 --warmup 3 --workload c2_1Mx768_b1024_k10_cosine_bruteforce --no-cpu-baseline --no-extras --no-parity > gpurun_out/c3_ncu_bench.log 2>&1; echo "rc=$?"
