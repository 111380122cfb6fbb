mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/step_times.py 2>&1 | grep -E "expand|total"
CF_TC_NEG=2 timeout 300 python tools/step_times.py 2>&1 | grep -E "total"
