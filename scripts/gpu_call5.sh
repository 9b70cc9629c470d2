#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== launch list multipass 1.25M"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c5_launches_multipass.csv python scripts/one_shard.py 1250000 multipass 5 > gpurun_out/c5_mp.log 2>&1; echo rc=$?
echo "== launch list stream 1.25M"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c5_launches_stream.csv python scripts/one_shard.py 1250000 stream 5 > gpurun_out/c5_st.log 2>&1; echo rc=$?
echo "== ncu full streaming kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_tc_kernel -s 7 -c 1 -o gpurun_out/c5_stream_full python scripts/one_shard.py 1250000 stream 5 > gpurun_out/c5_full.log 2>&1; echo rc=$?
echo "== ncu full multipass last pass"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_tc_kernel -s 29 -c 1 -o gpurun_out/c5_mp_full python scripts/one_shard.py 1250000 multipass 5 > gpurun_out/c5_full2.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
