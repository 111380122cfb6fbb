#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r2s_steps.log 2>&1; echo "steps rc=$?"; tail -1 gpurun_out/r2s_steps.log
cat gpurun_out/r2s_steps.log | awk -F'|' '{printf "%s|%s\n", $3,$5}' | tr -s ' ' | paste - - - - | head -14
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2s_bench.log",):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f).read()[-2000:])
PY
