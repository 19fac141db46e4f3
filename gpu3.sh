cd $GRAFT_REPO_ROOT
SWB_MAX_AHEAD=3 python scripts/profile_step.py config2_1GB 1 3000 > gpurun_out/plain.log 2>&1 && \
SWB_MAX_AHEAD=3 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:merge_rows -s 3000 -c 600 --csv --log-file gpurun_out/launches_loop.csv python scripts/profile_step.py config2_1GB 1 3000 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
