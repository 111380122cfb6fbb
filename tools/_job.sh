mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "taps or u8 or resize or batch32" 2>&1 | tail -2
timeout 300 python tools/step_times.py 2>&1 | sed -n 3,3p
