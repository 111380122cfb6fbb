#!/bin/bash
# One parameterised GPU job runner (replaces the per-session gpu_job_r*.sh scripts).
#   tools/gpu_job.sh <tag> <step> [<step> ...]
# steps: pytest | pytest:<expr> | steps[:ENV=V,...] | bench | bench_ref | smoke | heads_probe | sweep | ncu_list | ncu_full:<regex> |
#        sanitize:<tool> | sh:<command>
# every step writes gpurun_out/<tag>_<step>.log; a failing step does not stop the others.
set -u
tag=$1; shift
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for step in "$@"; do
  name=${step%%:*}; arg=""; [[ "$step" == *:* ]] && arg=${step#*:}
  safe=$(echo "$step" | tr -c 'A-Za-z0-9_.=,-' '_')
  log=gpurun_out/${tag}_${safe}.log
  echo "=== $step -> $log"
  case $name in
    pytest)      timeout 1500 python -m pytest tests -x -q -m gpu ${arg:+-k "$arg"} > $log 2>&1 ;;
    steps)       ( IFS=,; for kv in $arg; do export "$kv"; done; timeout 300 python tools/step_times.py ) > $log 2>&1 ;;
    bench)       timeout 600 python bench.py --steps 20 --warmup 5 > $log 2>&1 ;;
    bench_ref)   timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $log 2>&1 ;;
    smoke)       timeout 300 python __graft_entry__.py --smoke > $log 2>&1 ;;
    heads_probe) timeout 60 ./tools/build/heads_tc_probe 32 160 160 > $log 2>&1 ;;
    sweep)       timeout 900 python tools/size_sweep.py > $log 2>&1 ;;
    ncu_list)    timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 60 --csv \
                   --log-file gpurun_out/${tag}_launches.csv python tools/fwd_once.py --n 3 > $log 2>&1 ;;
    ncu_full)    timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$arg" -s 2 -c 3 -f -o gpurun_out/${tag}_full_$(echo "$arg" | tr -c 'A-Za-z0-9_' '_') \
                   python tools/fwd_once.py --n 2 > $log 2>&1 ;;
    sanitize)    timeout 1200 compute-sanitizer --tool $arg python tools/fwd_once.py --n 1 --batch 2 --size 320 > $log 2>&1 ;;
    sh)          timeout 1200 bash -c "$arg" > $log 2>&1 ;;
    *)           echo "unknown step $step" ;;
  esac
  echo "rc=$? ($step)"; tail -n 25 $log
done
