timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
timeout 120 python tools/mbf_trace.py --mask 0 --mbd 0x1 --j0 100 --nj 40 2>&1 | tail -22
