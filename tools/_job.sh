timeout 300 python tools/step_times.py 2>&1 | grep -E " pw " > gpurun_out/pw_base.log
CF_TC_DIRECT=1 timeout 300 python tools/step_times.py 2>&1 | grep -E " pw " > gpurun_out/pw_d1.log
CF_TC_DIRECT=2 timeout 300 python tools/step_times.py 2>&1 | grep -E " pw " > gpurun_out/pw_d2.log
paste gpurun_out/pw_base.log gpurun_out/pw_d1.log gpurun_out/pw_d2.log | awk -F'|' '{printf "%-36s %8s %8s %8s\n", $3, $5, $10, $15}'
