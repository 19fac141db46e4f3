# round 2, GPU call L (1 GPU): resident launch from a helper thread -- full GPU suite, smoke alone, smoke under ncu's launch list (the
# driver's own check: must complete now), config 2 + config 3 bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_l.log; echo "pytest done"; tail -4 gpurun_out/pytest_l.log )
( timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_l.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_l.log )
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/smoke_ncu_l.log 2>&1; echo "smoke under ncu rc=$?"; tail -2 gpurun_out/smoke_ncu_l.log )
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_smoke_launches.csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
  if len(r) > vi:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki].split("(")[0]; tot[k] += v; cnt[k] += 1
for k, v in tot.most_common(30): print(f"{k:40s} {cnt[k]:5d} launches {v/1e3:12.1f} us")
PY
( timeout 900 python bench.py --workload config2_1GB --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2l.json 2> gpurun_out/bench_c2l.log; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_c2l.log )
( timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3l.json 2> gpurun_out/bench_c3l.log; echo "bench c3 rc=$?"; grep -v "warmup" gpurun_out/bench_c3l.log | tail -2 )
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2l.json", "gpurun_out/bench_c3l.json"):
  try:
    d = json.load(open(f))
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2),
          "enc", {k: round(v) for k, v in d["extra"]["encode"].items() if k.endswith("MB_per_s")}, "parity", d["parity"]["equal"], "retried", len(d["retried_steps"]), "traffic", d["roofline"]["traffic"], d["roofline"]["actual_dram"])
  except Exception as e:
    print(f, "unreadable:", e)
PY
