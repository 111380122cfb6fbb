#!/bin/bash
mkdir -p gpurun_out
CF_PDL=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pw_tc -s 35 -c 11 -o gpurun_out/r2g_deep_pw -f python tools/fwd_once.py --n 2 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2g_ncu.log
CF_PDL=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_topk|k_heads|k_dwt" -s 14 -c 14 -o gpurun_out/r2g_misc -f python tools/fwd_once.py --n 2 > gpurun_out/r2g_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/*.ncu-rep
