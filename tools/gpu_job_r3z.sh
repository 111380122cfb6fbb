#!/bin/bash
# Refresh of the judged artefacts on the final build: bench line, per-launch times, ncu launch list + DRAM traffic.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r3z_bench.log 2>&1; echo "bench rc=$?"
timeout 200 python tools/step_times.py > gpurun_out/r3z_steps.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 129 -c 43 --csv --log-file gpurun_out/r3z_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3z_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r3z_bench.log") if l.startswith("{")][-1])
print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["clocks"], d.get("cpu_baseline",{}).get("value"))
print({k:v["ms"] for k,v in d["roofline"]["classes"].items()}, d["roofline"]["frac"], d["roofline"]["traffic"])
PY
tail -1 gpurun_out/r3z_steps.log
