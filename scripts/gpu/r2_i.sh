# round 2, GPU call I (1 GPU): timer-skew fix in; full suite; config 3 long run; ncu of the scripted merge_cluster launch without the cooperative / L2 attributes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; echo "pytest done"; tail -4 gpurun_out/pytest_gpu.log )
( timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3i.json 2> gpurun_out/bench_c3i.log; echo "bench c3 rc=$?"; grep -v "warmup" gpurun_out/bench_c3i.log | tail -3 )
python - <<'PY'
import json
for f in ("gpurun_out/bench_c3i.json",):
  try:
    d = json.load(open(f))
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2),
          "enc", {k: round(v) for k, v in d["extra"]["encode"].items() if k.endswith("MB_per_s")}, "parity", d["parity"]["equal"], "retried", len(d["retried_steps"]))
    print("   ", d.get("cpu_baseline", {}).get("ours_on_same_sample"))
  except Exception as e:
    print(f, "unreadable:", e)
PY
export SWB_NO_COOP=1 SWB_NO_L2_PERSIST=1
timeout 600 python scripts/profile_r2.py > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.log && cat gpurun_out/profile_plain.json &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:merge_cluster' -c 1 --csv --log-file gpurun_out/r2_merge_cluster_dram.csv python scripts/profile_r2.py > gpurun_out/ncu_m1.log 2>&1
echo "ncu dram rc=$?"; tail -3 gpurun_out/r2_merge_cluster_dram.csv | cut -c1-60,380-
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:merge_cluster' -c 1 -o gpurun_out/r2_merge_cluster python scripts/profile_r2.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
