CF_GRAPH=0 timeout 100 python tools/latency_b1.py 2>&1 | tail -3
timeout 100 python tools/latency_b1.py 2>&1 | tail -3
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
