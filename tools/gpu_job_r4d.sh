#!/bin/bash
# ncu --set full of fourteen stride-16/32 launches of the final build (b7 depth-wise ... b11 projection; launches 22-35 of the step).
mkdir -p gpurun_out
CF_PDL=0 timeout 115 ncu --set full --clock-control none -s 65 -c 14 -o gpurun_out/r4d_full_deep -f python tools/fwd_once.py --n 2 > gpurun_out/r4d_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r4d_ncu.log
timeout 30 python tools/ncu_full_summary.py gpurun_out/r4d_full_deep.ncu-rep "b7 dw5 s1 384ch" "b7 project 384->96" "b8 expand 96->576" "b8 dw5 s1 576ch" "b8 project 576->96 +res" "b9 expand 96->576" "b9 dw5 s2 576ch" "b9 project 576->160" "b10 expand 160->960" "b10 dw5 s1 960ch" "b10 project 960->160 +res" "b11 expand 160->960" "b11 dw3 s1 960ch" "b11 project 960->320" > gpurun_out/r4d_full_deep.md 2>&1
cat gpurun_out/r4d_full_deep.md | cut -c1-330
ls -la gpurun_out/*.ncu-rep
