#!/bin/bash
# GPU job: parity tests with PDL + 8x16 depth-wise tiles, per-launch times, bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r2b_steps.log 2>&1; echo "steps rc=$?"; tail -1 gpurun_out/r2b_steps.log
CF_DWT_GEOM=0 timeout 200 python tools/step_times.py > gpurun_out/r2b_steps_geom0.log 2>&1; tail -1 gpurun_out/r2b_steps_geom0.log
CF_DWT_GEOM=2 timeout 200 python tools/step_times.py > gpurun_out/r2b_steps_geom2.log 2>&1; tail -1 gpurun_out/r2b_steps_geom2.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_pdl1.log 2>&1; echo "bench rc=$?"
CF_PDL=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_pdl0.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2b_bench_pdl1.log","gpurun_out/r2b_bench_pdl0.log"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
