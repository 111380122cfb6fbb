#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3c_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r3c_steps.log 2>&1; echo "dw: $(grep '| dw' gpurun_out/r3c_steps.log | awk -F'|' '{printf "%s ", $5}')"; tail -1 gpurun_out/r3c_steps.log
for g in 0 1 2; do
CF_DWT_GEOM=$g timeout 200 python tools/step_times.py > gpurun_out/r3c_steps_g$g.log 2>&1; echo "geom $g: $(grep '| dw' gpurun_out/r3c_steps_g$g.log | awk -F'|' '{printf "%s ", $5}')"
done
