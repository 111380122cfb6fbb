#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r3j_tc_tune.jsonl
for shape in 960,320 960,160 576,160 576,96 384,96 384,64 192,64 320,24 160,960 96,576; do
  timeout 300 python tools/tc_tune.py --only $shape --out gpurun_out/r3j_tc_tune.jsonl > gpurun_out/r3j_tc_tune_${shape}.log 2>&1; echo "tune $shape rc=$?"
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r3j_tc_tune.jsonl')]
started=[r['key'] for r in rows if r.get('status')=='started']
done=[r['key'] for r in rows if 'rc' in r]
print('crashed:', [k for k in started if k not in done][:5])
rows=[r for r in rows if 'ms' in r]
for name in sorted(set(r['name'] for r in rows)):
    rs=sorted([r for r in rows if r['name']==name and r['status']=='ok'], key=lambda r:r['ms'])
    d=[r for r in rows if r['name']==name and not r['variant']]
    print(name, 'default', round(d[0]['ms']*1e3,1) if d else None)
    for r in rs[:2]: print('   ', round(r['ms']*1e3,1), r['plan'])
    for r in rows:
        if r['name']==name and 'mc=1' in r['plan']: print('   mc:', r['status'], round(r['ms']*1e3,1), r['plan'])
PY
