#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 3 4; do
  for m in 3 2; do
    CF_DWT_GEOM=$g CF_DWT_MIN2=$m timeout 200 python tools/step_times.py > gpurun_out/r2c_steps_g${g}_m${m}.log 2>&1
    echo "geom $g min2 $m: $(grep '| dw' gpurun_out/r2c_steps_g${g}_m${m}.log | awk -F'|' '{printf "%s ", $5}')"
  done
done
