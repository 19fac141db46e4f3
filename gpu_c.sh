cd $GRAFT_REPO_ROOT
SWB_TRACE_WAIT=1 timeout 120 python scripts/profile_step.py config2_1GB 1 > gpurun_out/trace0.log 2>&1
grep "trace\] hints" gpurun_out/trace0.log | tail -1 | cut -c1-500
free -g | head -2
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
if [ "$avail" -lt 48 ]; then echo "only $avail GB of host memory available: skipping the 10 GB check"; exit 0; fi
timeout 900 python scripts/scale_check.py config3_10GB 32768 2>&1 | tail -4 | cut -c1-1200
