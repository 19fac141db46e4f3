cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
timeout 300 python scripts/determinism_check.py config2_1GB 3 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_last.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['extra']['phase_ms'], d['extra']['us_per_merge'], d['extra']['host_split_ms'], d['extra']['look_ahead'])"
