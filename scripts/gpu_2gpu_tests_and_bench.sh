#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== all GPU tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/c16_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/c16_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench N=2 (p2p)"; timeout 600 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c16_bench_n2.json 2> gpurun_out/c16_bench_n2.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c16_bench_n2.json; tail -3 gpurun_out/c16_bench_n2.err
echo "== bench N=2 1M rows (small shards, p2p)"; timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --workload c2_1Mx768_b1024_k10_cosine_bruteforce > gpurun_out/c16_bench_c2_n2.json 2> gpurun_out/c16_bench_c2_n2.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c16_bench_c2_n2.json
echo "== bench N=1"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c16_bench_n1.json 2> gpurun_out/c16_bench_n1.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c16_bench_n1.json
