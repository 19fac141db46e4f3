cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-encode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step']); print(d['extra']['wall_ms']); print(d['extra']['e2e_wall_ms']); print(d['extra']['phase_ms'], d['extra']['e2e_phase_ms'])"
