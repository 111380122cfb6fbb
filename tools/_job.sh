timeout 120 python tools/mbf_check.py --mask 0x6 --mbd 0x1 --time 2>&1 | grep -E "block[0-2] |inds|MBF|fused|rror"
timeout 120 python tools/mbf_trace.py --mask 0x4 --j0 100 --nj 40 2>&1 | tail -21
