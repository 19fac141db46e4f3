cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -14; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Thread|Socket" 
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s4.json 2> gpurun_out/bench_s4.log; echo "bench rc=$?"
grep "bound" gpurun_out/bench_s4.log
python -c "
import json; d=json.load(open('gpurun_out/bench_s4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['extra']['phase_ms'], d['extra']['e2e_phase_ms'], d['extra']['us_per_merge']); r=d['roofline']; print(r['kernel'], r['achieved'], r['frac'], r['avg_launch_us'], r['resident_split']); print(d['extra']['encode']); print(d['extra']['look_ahead'], d['extra']['host_split_ms'], d['cpu_baseline']['value'], d['gpu_launches'])"
SWB_BENCH_NO_BIND=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nobind.json 2> gpurun_out/bench_nobind.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_nobind.json')); print('nobind', d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['us_per_merge'], d['extra']['encode']['e2e_MB_per_s'])"
