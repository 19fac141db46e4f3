cd $GRAFT_REPO_ROOT
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r1e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['extra']['phase_ms'], d['extra']['us_per_merge']); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['gpu_launches']); print(d['extra']['encode']); print(d['cpu_baseline']['value'])"
SWB_TRACE_WAIT=1 python scripts/profile_step.py config2_1GB 2 2>&1 | tail -8 | cut -c1-2500
