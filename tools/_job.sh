timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 100 python tools/step_times.py --iters 10 2>&1 | grep -E "heads|total us|fused|rror"
