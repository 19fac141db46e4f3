cd $GRAFT_REPO_ROOT
SWB_NO_PERSISTENT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_rows -s 3000 -c 3 -o gpurun_out/prof_merge_rows_r1b python scripts/profile_step.py config2_1GB 1 3200 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log | cut -c1-300
