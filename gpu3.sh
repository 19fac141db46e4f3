cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/profile_step.py config2_1GB 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print({k:round(v,4) for k,v in d.items() if k!='stats'})
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in s.items()})
"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-encode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench200', d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['phase_ms'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['extra']['ms_per_step_with_kernel_timing'])"
SWB_BENCH_CLOCK_MS=1000 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-encode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1000', d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['phase_ms'], d['clocks'])"
