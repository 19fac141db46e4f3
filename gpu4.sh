cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
SWB_DEVICE_LOOP=1 python -m pytest tests/test_gpu_train_parity.py -x -q -m gpu 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r1c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['phase_ms'], d['extra']['us_per_merge']); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_us'])"
