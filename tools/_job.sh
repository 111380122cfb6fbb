timeout 120 python tools/mbf_check.py --mask 0x6 --time 2>&1 | grep -E "block[1-3] |hm_sig|inds|MBF|fused|rror"
timeout 120 python tools/mbf_trace.py --mask 0x2 --j0 240 --nj 36 2>&1 | tail -19
