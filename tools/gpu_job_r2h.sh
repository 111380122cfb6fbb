#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2h_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r2h_steps.log 2>&1; echo "steps rc=$?"; tail -4 gpurun_out/r2h_steps.log
for shape in 384,96 576,96 96,576 960,160 960,320 576,160; do
for i in 1 2 3; do
  timeout 600 python tools/tc_tune.py --only $shape --out gpurun_out/r2h_tc_tune.jsonl > gpurun_out/r2h_tc_tune_${shape}_$i.log 2>&1
  rc=$?; echo "tune $shape pass $i rc=$rc"
  [ $rc -eq 0 ] && break
done
done
python - <<'PY'
import json
for line in open('gpurun_out/r2h_tc_tune.jsonl'):
    r=json.loads(line)
    if 'ms' not in r: continue
    if 'atmem=1' in r['plan'] or not r['variant']:
        print(f"{r['name']:8s} {r['status']:5s} {r['ms']*1e3:7.1f}  {r['plan']}")
PY
