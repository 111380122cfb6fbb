#!/bin/bash
# Final captures of the round: tests, the bench line, the reference arm, ncu launch list + DRAM traffic, ncu --set full of the top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z_bench.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/r2z_bench_ref.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
timeout 200 python tools/step_times.py > gpurun_out/r2z_steps.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 129 -c 43 --csv --log-file gpurun_out/r2z_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
CF_PDL=0 timeout 900 ncu --set full --import-source on --clock-control none -s 43 -c 12 -o gpurun_out/r2z_full_shallow -f python tools/fwd_once.py --n 2 > gpurun_out/r2z_ncu_full.log 2>&1; echo "ncu full rc=$?"
CF_PDL=0 timeout 600 ncu --set full --clock-control none -k regex:"k_heads|k_topk|k_peak" -s 3 -c 3 -o gpurun_out/r2z_full_tail -f python tools/fwd_once.py --n 2 > gpurun_out/r2z_ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2z_bench.log") if l.startswith("{")][-1])
print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["clocks"], d.get("cpu_baseline",{}).get("value"))
print({k:v["ms"] for k,v in d["roofline"]["classes"].items()})
print(d["roofline"].get("dominant_launch"))
PY
ls -la gpurun_out/*.ncu-rep
