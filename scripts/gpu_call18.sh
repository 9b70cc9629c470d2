#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SDB_TRACE=1 timeout 300 python scripts/trace_shard.py 1250000 2 14 2> gpurun_out/c18_trace_1250k_d2.txt; echo "rc=$?"
SDB_TRACE=1 timeout 300 python scripts/trace_shard.py 1250000 3 14 2> gpurun_out/c18_trace_1250k_d3.txt; echo "rc=$?"
grep -c TRACE gpurun_out/c18_trace_1250k_d2.txt
