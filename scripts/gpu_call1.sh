#!/bin/bash
# first GPU call of round 2: correctness of the streaming screen + async API, then bench A/B and a launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_gpu.txt 2>&1
nproc >> gpurun_out/c1_gpu.txt; free -g >> gpurun_out/c1_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?"
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_stream.py -q --timeout 300 -x > gpurun_out/c1_test_stream.log 2>&1; echo "stream rc=$?"; tail -5 gpurun_out/c1_test_stream.log
echo "== knn tests"; timeout 1200 python -m pytest tests/test_gpu_knn.py tests/test_gpu_multi.py -q --timeout 600 > gpurun_out/c1_test_knn.log 2>&1; echo "knn rc=$?"; tail -5 gpurun_out/c1_test_knn.log
echo "== bench streaming"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/c1_bench_stream.json 2> gpurun_out/c1_bench_stream.err; echo "rc=$?"; head -c 1500 gpurun_out/c1_bench_stream.json
echo "== bench multipass"; timeout 600 python bench.py --steps 20 --warmup 3 --schedule multipass --no-cpu-baseline --no-extras > gpurun_out/c1_bench_multipass.json 2> gpurun_out/c1_bench_multipass.err; echo "rc=$?"; head -c 600 gpurun_out/c1_bench_multipass.json
echo "== bench 1.25M shard sized"; timeout 600 python bench.py --steps 20 --warmup 3 --workload c2_1Mx768_b1024_k10_cosine_bruteforce --no-cpu-baseline --no-extras > gpurun_out/c1_bench_c2.json 2> gpurun_out/c1_bench_c2.err; echo "rc=$?"; head -c 600 gpurun_out/c1_bench_c2.json
echo "== bench clustered"; timeout 900 python bench.py --steps 10 --warmup 3 --data clustered --no-cpu-baseline --no-extras > gpurun_out/c1_bench_clustered.json 2> gpurun_out/c1_bench_clustered.err; echo "rc=$?"; head -c 1200 gpurun_out/c1_bench_clustered.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/c1_ncu_bench.log 2>&1; echo "rc=$?"
echo "== other gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_stream.py --deselect tests/test_gpu_knn.py --deselect tests/test_gpu_multi.py > gpurun_out/c1_test_rest.log 2>&1; echo "rest rc=$?"; tail -5 gpurun_out/c1_test_rest.log
