#!/bin/bash
mkdir -p gpurun_out
CF_PWN=0 timeout 200 python tools/step_times.py > gpurun_out/r3a_steps_pwn0.log 2>&1; grep "b0 project\|up[123]" gpurun_out/r3a_steps_pwn0.log; tail -1 gpurun_out/r3a_steps_pwn0.log
for e in 5 4; do
  timeout 200 python tools/step_times.py --pw $e > gpurun_out/r3a_steps_pw$e.log 2>&1; echo "engine $e: $(tail -1 gpurun_out/r3a_steps_pw$e.log)"
done
