mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/step_times.py 2>&1 | grep -E "res|total"
