#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
CF_STEM_TC=2 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2l_pytest_stem2.log 2>&1; echo "pytest(stem2) rc=$?"; tail -3 gpurun_out/r2l_pytest_stem2.log
timeout 200 python tools/step_times.py > gpurun_out/r2l_steps.log 2>&1; echo "steps rc=$?"; head -3 gpurun_out/r2l_steps.log | tail -1; tail -1 gpurun_out/r2l_steps.log
CF_STEM_TC=2 timeout 200 python tools/step_times.py > gpurun_out/r2l_steps_stem2.log 2>&1; echo "steps rc=$?"; head -3 gpurun_out/r2l_steps_stem2.log | tail -1; tail -1 gpurun_out/r2l_steps_stem2.log
