#!/bin/bash
mkdir -p gpurun_out
for n in 1 3 5 6; do
CF_PWN_NKB=$n timeout 200 python tools/step_times.py > gpurun_out/r3m_steps_nkb$n.log 2>&1; echo "nkb<=$n: $(grep 'project' gpurun_out/r3m_steps_nkb$n.log | head -5 | awk -F'|' '{printf "%s=%s ", $3,$5}') $(tail -1 gpurun_out/r3m_steps_nkb$n.log | cut -c1-20)"
done
CF_PWN_NKB=6 CF_PWN_CTAS=2 timeout 200 python tools/step_times.py > gpurun_out/r3m_steps_nkb6c2.log 2>&1; echo "nkb<=6 ctas2: $(grep 'project' gpurun_out/r3m_steps_nkb6c2.log | head -5 | awk -F'|' '{printf "%s=%s ", $3,$5}')"
