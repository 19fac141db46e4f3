# round 2, GPU call Q (1 GPU): hints after GRID merges -- full GPU suite, config 2 + config 3 bench, pre-pass throughput
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_q.log; echo "pytest done"; tail -4 gpurun_out/pytest_q.log )
( timeout 900 python bench.py --workload config2_1GB --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2q.json 2> gpurun_out/bench_c2q.log; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_c2q.log | cut -c1-300 )
( timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3q.json 2> gpurun_out/bench_c3q.log; echo "bench c3 rc=$?"; grep -v "warmup" gpurun_out/bench_c3q.log | tail -2 | cut -c1-300 )
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2q.json", "gpurun_out/bench_c3q.json"):
  try:
    d = json.load(open(f))
    rs = d["roofline"]["resident_split"]
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2),
          "dev us", round(d["roofline"]["avg_launch_us"], 2), "grid", rs["grid_merges"], round(rs["grid_us_per_merge"], 1), "local", round(rs["local_us_per_merge"], 2),
          d["extra"]["look_ahead"], d["extra"]["host_split_ms"], "parity", d["parity"]["equal"], "retried", len(d["retried_steps"]))
  except Exception as e:
    print(f, "unreadable:", e)
PY
( timeout 300 python scripts/profile_prepass.py 1000000000 > gpurun_out/prepass_q.json 2> gpurun_out/prepass_q.log; cat gpurun_out/prepass_q.json )
