mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.log 2>&1; echo bench rc=$?
timeout 300 python tools/step_times.py > gpurun_out/r2h_steps.log 2>&1; echo steps rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 160 -c 40 --csv --log-file gpurun_out/r2h_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_traffic_bench.log 2>&1; echo traffic rc=$?
timeout 300 python tools/mbf_trace.py --mask 0x4 --j0 200 --nj 60 > gpurun_out/trace_b2.log 2>&1
timeout 300 python tools/mbf_trace.py --mask 0x0 --mbd 0x1 --j0 200 --nj 60 > gpurun_out/trace_l0.log 2>&1
timeout 300 python tools/mbf_trace.py --mask 0x2 --j0 200 --nj 60 > gpurun_out/trace_b1.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/fwd_once.py --n 1 --batch 2 --size 320 > gpurun_out/r2h_memcheck.log 2>&1; echo memcheck rc=$?; tail -1 gpurun_out/r2h_memcheck.log
