#!/bin/bash
# 8-GPU box: north-star scaling point, C4 sharded, C5 graph sharded.  Charged 8x: keep it short.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== northstar N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c8_bench_n8.json 2> gpurun_out/c8_bench_n8.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c8_bench_n8.json
echo "== northstar N=4"; timeout 600 $TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c8_bench_n4.json 2> gpurun_out/c8_bench_n4.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c8_bench_n4.json
echo "== C4 N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload c4_10Mx1536_b4096_k100_cosine_bruteforce > gpurun_out/c8_bench_c4_n8.json 2> gpurun_out/c8_bench_c4_n8.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c8_bench_c4_n8.json
echo "== C5 graph N=8 (4096 sources)"; timeout 900 $TR --nproc-per-node 8 --master-port 29514 bench_extra.py graph --nodes 50000000 --edges 500000000 --sources 4096 --collect-sources 64 --checksum --no-cpu > gpurun_out/c8_graph_c5_n8.json 2> gpurun_out/c8_graph_c5_n8.err; echo "rc=$?"; cut -c1-700 gpurun_out/c8_graph_c5_n8.json
for f in gpurun_out/c8_*.err; do echo "--- $f"; tail -n 3 $f; done
