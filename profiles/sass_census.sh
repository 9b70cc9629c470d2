#!/bin/bash
# opcode census of the shipped library (run from the repo root): proves the contraction kernel is tcgen05 / TMEM / TMA native
so=surrealdb_b200/csrc/libsdbgpu.so
echo "# cuobjdump -sass $so | opcode histogram (selected: tensor core, TMEM, TMA, bulk copy, barriers, FP64, atomics)"
cuobjdump -sass $so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn > /tmp/sass_all.txt
grep -E " (UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|UTCATOMSWS|UTMALDG|UTMACCTL|UBLKCP|SYNCS|REDUX|CREDUX|ATOMG|ATOMS|REDG|DFMA|DADD|DMUL|NANOSLEEP|CS2R|MATCH|VOTE|SHFL)$" /tmp/sass_all.txt
echo "# total distinct opcodes: $(wc -l < /tmp/sass_all.txt), total instructions: $(awk '{s+=$1} END {print s}' /tmp/sass_all.txt)"
