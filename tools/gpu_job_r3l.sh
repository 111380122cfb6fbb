#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/step_times.py > gpurun_out/r3l_steps.log 2>&1; grep "b0 project\|up[23]" gpurun_out/r3l_steps.log; tail -1 gpurun_out/r3l_steps.log
CF_PWN_CTAS=3 timeout 200 python tools/step_times.py > gpurun_out/r3l_steps_pwn3.log 2>&1; grep "b0 project\|up[23]" gpurun_out/r3l_steps_pwn3.log; tail -1 gpurun_out/r3l_steps_pwn3.log
