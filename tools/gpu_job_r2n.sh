#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 9 2 4 13 15; do
  rm -f gpurun_out/r2n_dbg_$dbg.jsonl
  CF_TC_DEBUG=$dbg timeout 300 python tools/tc_tune.py --default-only --out gpurun_out/r2n_dbg_$dbg.jsonl 2>&1 | grep -E "^(b0|b1|b2|b3|b4)\.\S+\s+M=" | awk -v d=$dbg '{printf "dbg=%s %s K=%s N=%s %s us | ", d, $1, $5, $7, $9}'
  echo
done
