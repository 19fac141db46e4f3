# round 2, GPU call N (1 GPU): smoke under ncu's launch list with no switches (the driver's check); ncu --set full of the pre-pass kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/smoke_ncu_n.log 2>&1; echo "smoke under ncu (no switches): rc=$?"; tail -1 gpurun_out/smoke_ncu_n.log; grep "merge_cluster" gpurun_out/r2_smoke_launches.csv | rev | cut -d, -f1-3 | rev )
( timeout 300 python scripts/profile_prepass.py > gpurun_out/prepass_plain.json 2> gpurun_out/prepass_plain.log; cat gpurun_out/prepass_plain.json )
( timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:stream_map' -c 2 -o gpurun_out/r2_prepass python scripts/profile_prepass.py > gpurun_out/ncu_prepass.log 2>&1; echo "ncu prepass rc=$?"; tail -2 gpurun_out/ncu_prepass.log )
