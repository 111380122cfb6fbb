#!/bin/bash
mkdir -p gpurun_out
CF_STEM_TC=2 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest_stem2.log 2>&1; echo "pytest(stem2) rc=$?"; tail -8 gpurun_out/r2k_pytest_stem2.log
CF_STEM_TC=2 timeout 200 python tools/step_times.py > gpurun_out/r2k_steps_stem2.log 2>&1; echo "steps rc=$?"; head -4 gpurun_out/r2k_steps_stem2.log | tail -2; tail -1 gpurun_out/r2k_steps_stem2.log
timeout 200 python tools/step_times.py > gpurun_out/r2k_steps.log 2>&1; echo "steps rc=$?"; grep "| dw" gpurun_out/r2k_steps.log | awk -F'|' '{printf "%s ", $5}'; echo; tail -1 gpurun_out/r2k_steps.log
