#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2e_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r2e_steps.log 2>&1; echo "steps rc=$?"; tail -1 gpurun_out/r2e_steps.log
CF_TC_RCHUNK=0 timeout 200 python tools/step_times.py > gpurun_out/r2e_steps_rc0.log 2>&1; tail -1 gpurun_out/r2e_steps_rc0.log
for i in 1 2 3 4 5 6; do
  timeout 900 python tools/tc_tune.py --out gpurun_out/r2e_tc_tune.jsonl > gpurun_out/r2e_tc_tune_$i.log 2>&1
  rc=$?; echo "tune pass $i rc=$rc"
  [ $rc -eq 0 ] && break
done
sed -n '/| layer/,$p' gpurun_out/r2e_tc_tune_*.log | tail -30
