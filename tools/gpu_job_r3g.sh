#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r3g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r3g_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python __graft_entry__.py --smoke > gpurun_out/r3g_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r3g_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 10 python __graft_entry__.py --smoke > gpurun_out/r3g_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r3g_synccheck.log
