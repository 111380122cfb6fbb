#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r2q_steps.log 2>&1; echo "steps rc=$?"; tail -1 gpurun_out/r2q_steps.log
cat gpurun_out/r2q_steps.log | awk -F'|' '{printf "%s|%s\n", $3,$5}' | tr -s ' ' | paste - - - - | head -14
