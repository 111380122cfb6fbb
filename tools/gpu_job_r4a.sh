#!/bin/bash
# Final confirmation of the parity suite + the size / batch sweep (configs[3], configs[4] per-GPU shards, F5-derived decode timings).
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/r4a_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r4a_pytest.log
timeout 120 python tools/size_sweep.py --steps 10 --out gpurun_out/r4a_size_sweep.md > gpurun_out/r4a_size_sweep.log 2>&1; echo "sweep rc=$?"
tail -25 gpurun_out/r4a_size_sweep.log | cut -c1-400
