#!/bin/bash
mkdir -p gpurun_out
for shape in 96,576 960,320 576,96 24,144 160,960; do
  for dbg in 0 1 2 4 3 6 7; do
    CF_TC_DEBUG=$dbg timeout 300 python tools/tc_tune.py --only $shape --out gpurun_out/r2f_dbg_${shape}_${dbg}.jsonl 2>&1 | grep -E "^\S+\s+M=" | head -1 | sed "s/^/dbg=$dbg /"
  done
done
