cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1a.log
python scripts/profile_step.py config2_1GB 1 1200 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:merge_rows -s 1000 -c 3 -o gpurun_out/prof_merge_rows_r1 python scripts/profile_step.py config2_1GB 1 1200 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
