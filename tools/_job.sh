mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/fwd_once.py --n 1 --batch 1 --size 160 > gpurun_out/r2f_racecheck_analysis.log 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/r2f_racecheck_analysis.log
timeout 900 compute-sanitizer --tool synccheck python tools/fwd_once.py --n 1 --batch 2 --size 320 > gpurun_out/r2f_synccheck.log 2>&1; echo synccheck rc=$?; grep -c "Barrier error" gpurun_out/r2f_synccheck.log; grep -m3 -A6 "Barrier error" gpurun_out/r2f_synccheck.log | head -30; tail -3 gpurun_out/r2f_synccheck.log
