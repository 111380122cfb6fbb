mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/step_times.py 2>&1 | grep -E "b3 expand|b4 expand|b5 expand|total"
