mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; tail -2 gpurun_out/r2k_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2k_smoke.log 2>&1; tail -1 gpurun_out/r2k_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.log 2>&1; echo bench rc=$?
timeout 300 python tools/step_times.py > gpurun_out/r2k_steps.log 2>&1; echo steps rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 160 -c 40 --csv --log-file gpurun_out/r2k_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_traffic_bench.log 2>&1; echo traffic rc=$?
timeout 1200 ncu --set full --clock-control none --import-source on -s 38 -c 38 -f -o gpurun_out/r2_all python tools/fwd_once.py --n 2 > gpurun_out/r2_all.log 2>&1; echo ncu_all rc=$?
timeout 600 compute-sanitizer --tool memcheck python tools/fwd_once.py --n 1 --batch 2 --size 320 > gpurun_out/r2k_memcheck.log 2>&1; tail -1 gpurun_out/r2k_memcheck.log
timeout 900 compute-sanitizer --tool synccheck python tools/fwd_once.py --n 1 --batch 2 --size 320 > gpurun_out/r2k_synccheck.log 2>&1; tail -1 gpurun_out/r2k_synccheck.log
