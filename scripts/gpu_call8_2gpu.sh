#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c8_gpus.txt
echo "== 2-GPU tests"; timeout 900 python -m pytest tests/test_gpu_stream.py -q --timeout 300 -k "two_gpus or one_process or sharded" > gpurun_out/c8_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/c8_tests.log
echo "== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c8_bench_n1.json 2> gpurun_out/c8_bench_n1.err; echo "rc=$?"; head -c 250 gpurun_out/c8_bench_n1.json; echo
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err; echo "rc=$?"; head -c 400 gpurun_out/c8_bench_n2.json; echo; tail -5 gpurun_out/c8_bench_n2.err
echo "== graph 2^24 1 GPU"; timeout 600 python bench_extra.py graph --log2-nodes 24 --edges 160000000 --checksum > gpurun_out/c8_graph_1gpu.json 2> gpurun_out/c8_graph_1gpu.err; echo "rc=$?"; head -c 600 gpurun_out/c8_graph_1gpu.json; echo
echo "== graph 2^24 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_extra.py graph --log2-nodes 24 --edges 160000000 --checksum --no-cpu > gpurun_out/c8_graph_2gpu.json 2> gpurun_out/c8_graph_2gpu.err; echo "rc=$?"; head -c 600 gpurun_out/c8_graph_2gpu.json; echo; tail -5 gpurun_out/c8_graph_2gpu.err
