# scratch job of the last profiling call (tools/gpu_job.sh is the parameterised runner): DRAM traffic of one bench step and one
# `--set full` row per launch, summarised ON the box (the report itself is ~70 MB; gpurun_out/ travels back up to 64 MB)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 140 -c 90 --csv --log-file gpurun_out/r3_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3_traffic_bench.log 2>&1; echo traffic rc=$?
timeout 1200 ncu --set full --clock-control none -s 37 -c 37 -f -o /tmp/r3_all python tools/fwd_once.py --n 2 > gpurun_out/r3_all.log 2>&1; echo ncu_all rc=$?
python tools/ncu_full_summary.py /tmp/r3_all.ncu-rep > gpurun_out/r3_ncu_rows.md 2> gpurun_out/r3_ncu_rows.err; echo summary rc=$?
