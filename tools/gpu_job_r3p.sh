#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/step_times.py > gpurun_out/r3p_steps.log 2>&1; echo "default: $(grep 'b0 project\|b1 project\|up[123]' gpurun_out/r3p_steps.log | awk -F'|' '{printf "%s=%s ", $3,$5}') $(tail -1 gpurun_out/r3p_steps.log | cut -c1-20)"
CF_PWN_CTAS=3 timeout 200 python tools/step_times.py > gpurun_out/r3p_steps_c3.log 2>&1; echo "ctas3: $(grep 'b0 project\|b1 project\|up[123]' gpurun_out/r3p_steps_c3.log | awk -F'|' '{printf "%s=%s ", $3,$5}') $(tail -1 gpurun_out/r3p_steps_c3.log | cut -c1-20)"
