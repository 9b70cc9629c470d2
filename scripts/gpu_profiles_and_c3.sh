#!/bin/bash
# round-2 profile captures (1 GPU): launch list of the bench command, full capture of the streaming screen at 10M rows,
# full capture of the HNSW walk.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/c17_launch_bench.log 2>&1; echo "rc=$?"
echo "== full capture: streaming screen at 10M (probe + main launches of one batch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_tc_kernel --launch-skip 4 --launch-count 2 -f -o gpurun_out/r2_screen_stream_int8 python scripts/one_shard.py 10000000 stream 4 > gpurun_out/c17_ncu_screen.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/c17_ncu_screen.log
echo "== full capture: HNSW walk"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:hnsw_search -c 1 -f -o gpurun_out/r2_hnsw_walk python tests/dev/hnsw_ncu.py 1000000 > gpurun_out/c17_ncu_hnsw.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/c17_ncu_hnsw.log
echo "== C3: 10M x 768 HNSW"
timeout 1500 python bench_extra.py hnsw --rows 10000000 --dim 768 --metric cosine --builder incremental --efc 150 --queries 20000 > gpurun_out/c17_hnsw_c3_10M.json 2> gpurun_out/c17_hnsw_c3_10M.err; echo "rc=$?"; tail -4 gpurun_out/c17_hnsw_c3_10M.err; cut -c1-1500 gpurun_out/c17_hnsw_c3_10M.json
echo "== clustered data, 1 GPU"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --data clustered > gpurun_out/c17_bench_clustered.json 2> gpurun_out/c17_bench_clustered.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c17_bench_clustered.json
echo "== C4 on 1 GPU"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --workload c4_10Mx1536_b4096_k100_cosine_bruteforce > gpurun_out/c17_bench_c4_1gpu.json 2> gpurun_out/c17_bench_c4_1gpu.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c17_bench_c4_1gpu.json
echo "== headline with cpu baseline + extras"; timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/c17_bench_full.json 2> gpurun_out/c17_bench_full.err; echo "rc=$?"; python scripts/show_bench.py gpurun_out/c17_bench_full.json
ls -la gpurun_out/*.ncu-rep
