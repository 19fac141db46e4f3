cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 > gpurun_out/bench1.json 2> gpurun_out/bench1.log; echo "bench rc=$?"
tail -20 gpurun_out/bench1.log; cat gpurun_out/bench1.json
