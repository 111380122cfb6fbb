#!/bin/bash
# Re-run of the size / batch sweep with at most 4 rotating input batches (launch-plan cache hits, as in bench.py).
mkdir -p gpurun_out
timeout 120 python tools/size_sweep.py --steps 10 --out gpurun_out/r4c_size_sweep.md > gpurun_out/r4c_size_sweep.log 2>&1; echo "sweep rc=$?"
tail -25 gpurun_out/r4c_size_sweep.log | cut -c1-400
