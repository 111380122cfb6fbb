mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.log 2>&1; echo bench rc=$?
timeout 300 python tools/step_times.py > gpurun_out/r2f_steps.log 2>&1; echo steps rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 160 -c 40 --csv --log-file gpurun_out/r2f_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_traffic_bench.log 2>&1; echo traffic rc=$?
timeout 600 compute-sanitizer --tool memcheck python tools/fwd_once.py --n 1 --batch 2 --size 320 > gpurun_out/r2f_memcheck.log 2>&1; echo memcheck rc=$?; tail -2 gpurun_out/r2f_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/fwd_once.py --n 1 --batch 1 --size 160 > gpurun_out/r2f_racecheck.log 2>&1; echo racecheck rc=$?; tail -2 gpurun_out/r2f_racecheck.log
