mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 160 -c 40 --csv --log-file gpurun_out/r2_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_traffic_bench.log 2>&1; echo traffic rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mbf|k_heads_tc" -s 3 -c 3 -f -o gpurun_out/r2_full_fused python tools/fwd_once.py --n 2 > gpurun_out/r2_full.log 2>&1; echo full rc=$?
timeout 600 compute-sanitizer --tool memcheck python tools/fwd_once.py --n 1 --batch 2 --size 320 > gpurun_out/r2_memcheck.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/r2_memcheck.log
timeout 900 compute-sanitizer --tool synccheck python tools/fwd_once.py --n 1 --batch 2 --size 320 > gpurun_out/r2_synccheck.log 2>&1; echo synccheck rc=$?; tail -3 gpurun_out/r2_synccheck.log
