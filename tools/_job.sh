mkdir -p gpurun_out
timeout 300 python tools/mbf_check.py --mask 0x6 --mbd 0x1 --time 2>&1 | tail -4
timeout 300 python tools/mbf_trace.py --mask 0x4 --j0 200 --nj 60 > gpurun_out/trace_b2.log 2>&1
