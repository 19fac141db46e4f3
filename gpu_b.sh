cd $GRAFT_REPO_ROOT
for nb in 0; do
SWB_NO_BIRTH_LOG=$nb SWB_TRACE_WAIT=1 python scripts/profile_step.py config2_1GB 1 > gpurun_out/trace$nb.log 2>&1
grep "trace" gpurun_out/trace$nb.log | cut -c1-400
done
