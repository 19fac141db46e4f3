cd $GRAFT_REPO_ROOT
timeout 300 python scripts/determinism_check.py config2_1GB 3 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-encode > gpurun_out/bench_rep$i.json 2> gpurun_out/bench_rep$i.log
python -c "
import json; d=json.load(open('gpurun_out/bench_rep$i.json')); print($i, round(d['value'],3), round(d['ms_per_step'],1), round(d['e2e']['value'],3), round(d['e2e']['ms_per_step'],1), d['extra']['phase_ms'], round(d['extra']['us_per_merge'],2), round(d['roofline']['avg_launch_us'],2), round(d['roofline']['frac'],3))"
done
timeout 900 python scripts/scale_check.py config3_10GB 32768 2>&1 | tail -1 | cut -c1-700
