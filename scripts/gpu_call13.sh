#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SDB_HNSW_OCC=4
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:hnsw_search -c 1 -f -o gpurun_out/r2_hnsw_walk python tests/dev/hnsw_ncu.py 1000000 > gpurun_out/c13_ncu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c13_ncu.log
ls -la gpurun_out/r2_hnsw_walk.ncu-rep
