#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== northstar N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c8b_bench_n8.json 2> gpurun_out/c8b_bench_n8.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c8b_bench_n8.json
echo "== northstar N=8 (again, 50 steps)"; timeout 600 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/c8b_bench_n8_50.json 2> gpurun_out/c8b_bench_n8_50.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c8b_bench_n8_50.json
echo "== northstar N=4"; timeout 600 $TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c8b_bench_n4.json 2> gpurun_out/c8b_bench_n4.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c8b_bench_n4.json
echo "== C4 N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload c4_10Mx1536_b4096_k100_cosine_bruteforce > gpurun_out/c8b_bench_c4_n8.json 2> gpurun_out/c8b_bench_c4_n8.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c8b_bench_c4_n8.json
for f in gpurun_out/c8b_*.err; do echo "--- $f"; tail -n 3 $f; done
