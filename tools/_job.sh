mkdir -p gpurun_out
timeout 300 python tools/mbf_check.py --mask 0x6 --mbd 0x1 --time 2>&1 | tail -4
