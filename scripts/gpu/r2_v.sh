# round 2, GPU call V (1 GPU): ncu launch list of the bench command itself (config 3 default and config 2), after the same command has exited 0 without ncu
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for W in config3_10GB config2_1GB; do
  ( timeout 900 python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-encode > gpurun_out/bench_plain_$W.json 2> gpurun_out/bench_plain_$W.log; echo "bench $W plain rc=$?" )
  ( timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench_$W.csv python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-encode > gpurun_out/bench_ncu_$W.json 2> gpurun_out/bench_ncu_$W.log; echo "bench $W under ncu rc=$?" )
  python - <<PY
import csv, collections, json
rows = list(csv.reader(open("gpurun_out/r2_launches_bench_$W.csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
  if len(r) > vi:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki].split("(")[0][:60]; tot[k] += v; cnt[k] += 1
T = sum(tot.values())
for k, v in tot.most_common(8): print(f"  {k:60s} {cnt[k]:5d} launches {v/1e6:10.2f} ms  share {v/T:.3f}")
try:
  d = json.load(open("gpurun_out/bench_plain_$W.json")); print("  plain:", round(d["ms_per_step"], 1), "ms/step, kernel_share_of_step", round(d["roofline"]["kernel_share_of_step"], 3))
except Exception as e: print("  plain unreadable", e)
PY
done
