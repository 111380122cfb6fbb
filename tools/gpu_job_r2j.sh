#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2j_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r2j_steps.log 2>&1; echo "steps rc=$?"; grep "up[123]" gpurun_out/r2j_steps.log; tail -1 gpurun_out/r2j_steps.log
CF_DWT_DEBUG=1 timeout 200 python tools/step_times.py > gpurun_out/r2j_steps_dbg.log 2>&1
echo "dw tma-only: $(grep '| dw' gpurun_out/r2j_steps_dbg.log | awk -F'|' '{printf "%s ", $5}')"
for g in 0 1 2 3 4; do
  for m in 3 2; do
    CF_DWT_GEOM=$g CF_DWT_MIN2=$m timeout 200 python tools/step_times.py > gpurun_out/r2j_steps_g${g}_m${m}.log 2>&1
    echo "geom $g min2 $m: $(grep '| dw' gpurun_out/r2j_steps_g${g}_m${m}.log | awk -F'|' '{printf "%s ", $5}')"
  done
done
