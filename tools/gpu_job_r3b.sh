#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3b_pytest.log
timeout 200 python tools/step_times.py > gpurun_out/r3b_steps.log 2>&1; grep "up[123]" gpurun_out/r3b_steps.log; tail -1 gpurun_out/r3b_steps.log
rm -f gpurun_out/r3b_tc_tune.jsonl
for shape in 16,96 24,144 32,192 64,384 96,576 160,960; do
  timeout 600 python tools/tc_tune.py --only $shape --out gpurun_out/r3b_tc_tune.jsonl > gpurun_out/r3b_tc_tune_${shape}.log 2>&1; echo "tune $shape rc=$?"
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r3b_tc_tune.jsonl')]
rows=[r for r in rows if 'ms' in r]
for name in sorted(set(r['name'] for r in rows)):
    rs=sorted([r for r in rows if r['name']==name and r['status']=='ok'], key=lambda r:r['ms'])
    d=[r for r in rows if r['name']==name and not r['variant']]
    print(name, 'default', round(d[0]['ms']*1e3,1) if d else None)
    for r in rs[:4]: print('   ', round(r['ms']*1e3,1), r['plan'], r['variant'].get('CF_TC_STG'))
    bad=[r for r in rows if r['name']==name and r['status']!='ok']
    if bad: print('   NOT OK:', len(bad), bad[0]['plan'], bad[0]['variant'])
PY
