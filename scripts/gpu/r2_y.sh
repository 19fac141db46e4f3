# round 2, GPU call Y (1 GPU): the committed final build -- full GPU suite, smoke, one short config 3 bench with the encoder
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_y.log; echo "pytest done"; tail -3 gpurun_out/pytest_y.log )
( timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_y.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_y.log )
( timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3y.json 2> gpurun_out/bench_c3y.log; echo "bench c3 rc=$?" )
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c3y.json")); e = d["extra"]
print("value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), e["phase_ms"], "dev us", round(d["roofline"]["avg_launch_us"], 2), e["look_ahead"], "enc", {k: round(v) for k, v in e["encode"].items() if k.endswith("MB_per_s")}, "parity", d["parity"]["equal"], "retried", len(d["retried_steps"]))
PY
