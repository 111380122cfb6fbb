#!/bin/bash
# GPU job: baseline per-launch times, GEMM plan sweep, depth-wise occupancy probe.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pw_gemm.py -x -q -m gpu > gpurun_out/r2a_pwtest.log 2>&1; echo "pwtest rc=$?"
timeout 200 python tools/step_times.py > gpurun_out/r2a_steps_base.log 2>&1; echo "steps rc=$?"
for i in 1 2 3 4 5 6; do
  timeout 900 python tools/tc_tune.py --out gpurun_out/r2a_tc_tune.jsonl > gpurun_out/r2a_tc_tune_$i.log 2>&1
  rc=$?; echo "tune pass $i rc=$rc"
  [ $rc -eq 0 ] && break
done
CF_DWT_CTAS=1 timeout 200 python tools/step_times.py > gpurun_out/r2a_steps_dwt1.log 2>&1; echo "dwt1 rc=$?"
CF_DWT_NST=3 timeout 200 python tools/step_times.py > gpurun_out/r2a_steps_nst3.log 2>&1; echo "nst3 rc=$?"
tail -3 gpurun_out/r2a_steps_base.log
