cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s4.json 2> gpurun_out/bench_s4.log; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_s4.json
