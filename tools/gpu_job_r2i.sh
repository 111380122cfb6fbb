#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2i_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r2i_steps.log 2>&1; echo "steps rc=$?"; grep "| dw" gpurun_out/r2i_steps.log | awk -F'|' '{printf "%s ", $5}'; echo; tail -1 gpurun_out/r2i_steps.log
for i in 1 2 3 4 5 6; do
  timeout 900 python tools/tc_tune.py --out gpurun_out/r2i_tc_tune.jsonl > gpurun_out/r2i_tc_tune_$i.log 2>&1
  rc=$?; echo "tune pass $i rc=$rc"
  [ $rc -eq 0 ] && break
done
sed -n '/| layer/,$p' gpurun_out/r2i_tc_tune_*.log | tail -30
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2i_bench.log",):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f).read()[-2000:])
PY
