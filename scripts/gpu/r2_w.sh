# round 2, GPU call W (1 GPU): look-ahead list reused across merges -- full GPU suite, then config 3 with SWB_PEEK_DEPTH = 6 (default) / 2 (as before) / 4 / 8, config 2 with the default
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_w.log; echo "pytest done"; tail -3 gpurun_out/pytest_w.log )
for D in 6 2 4 8; do
  ( SWB_PEEK_DEPTH=$D timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-encode > gpurun_out/bench_c3w_$D.json 2> gpurun_out/bench_c3w_$D.log; echo "bench c3 depth $D rc=$?" )
done
( timeout 900 python bench.py --workload config2_1GB --steps 5 --warmup 3 --no-cpu-baseline --no-encode > gpurun_out/bench_c2w.json 2> gpurun_out/bench_c2w.log; echo "bench c2 rc=$?" )
python - <<'PY'
import json
for f in ["gpurun_out/bench_c3w_%d.json" % d for d in (6, 2, 4, 8)] + ["gpurun_out/bench_c2w.json"]:
  try:
    d = json.load(open(f)); e = d["extra"]
    print(f, "ms", round(d["ms_per_step"], 1), "merge_ms", round(e["phase_ms"]["merge_ms"], 1), "us/merge", round(e["us_per_merge"], 2), "dev us", round(d["roofline"]["avg_launch_us"], 2),
          e["look_ahead"], e["host_split_ms"], "parity", d["parity"]["equal"], "retried", len(d["retried_steps"]))
  except Exception as ex:
    print(f, "unreadable:", ex)
PY
