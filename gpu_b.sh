cd $GRAFT_REPO_ROOT
SWB_TRACE_WAIT=1 timeout 120 python scripts/profile_step.py config2_1GB 1 > gpurun_out/trace0.log 2>&1
grep "trace\|rror" gpurun_out/trace0.log | tail -3 | cut -c1-500
