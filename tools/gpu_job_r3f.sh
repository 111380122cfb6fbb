#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r3f_tc_tune.jsonl
for shape in 24,144 32,192 64,384 192,64 384,64; do
  timeout 600 python tools/tc_tune.py --only $shape --out gpurun_out/r3f_tc_tune.jsonl > gpurun_out/r3f_tc_tune_${shape}.log 2>&1; echo "tune $shape rc=$?"
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r3f_tc_tune.jsonl')]
rows=[r for r in rows if 'ms' in r]
for name in sorted(set(r['name'] for r in rows)):
    rs=sorted([r for r in rows if r['name']==name and r['status']=='ok'], key=lambda r:r['ms'])
    d=[r for r in rows if r['name']==name and not r['variant']]
    print(name, 'default', round(d[0]['ms']*1e3,1) if d else None)
    for r in rs[:3]: print('   ', round(r['ms']*1e3,1), r['plan'])
    for r in rows:
        if r['name']==name and 'nacc=3' in r['plan']: print('   nacc3:', r['status'], round(r['ms']*1e3,1), r['plan'])
PY
