#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2r_tc_tune.jsonl
for i in 1 2 3 4 5 6; do
  timeout 900 python tools/tc_tune.py --out gpurun_out/r2r_tc_tune.jsonl > gpurun_out/r2r_tc_tune_$i.log 2>&1
  rc=$?; echo "tune pass $i rc=$rc"
  [ $rc -eq 0 ] && break
done
sed -n '/| layer/,$p' gpurun_out/r2r_tc_tune_*.log | tail -30
