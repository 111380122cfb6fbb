for m in 0x6; do echo "== mask $m"; timeout 120 python tools/mbf_check.py --mask $m --time 2>&1 | grep -E "block[1-3] |hm_sig|inds|MBF|fused|rror"; done
for d in 3; do echo "== dbg $d"; CF_MBF_DEBUG=$d CF_MBF=0x6 timeout 60 python tools/step_times.py --iters 5 2>&1 | grep -E "fused|rror"; done
