# round 2, GPU call H (1 GPU): SOLO threshold 1024, encoder without the compaction buffer; single-pass ncu metrics of the scripted merge_cluster launch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_encode_parity.py tests/test_gpu_fuzz.py tests/test_gpu_variants.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/pytest_h.log; tail -4 gpurun_out/pytest_h.log )
( timeout 900 python bench.py --workload config2_1GB --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2h.json 2> gpurun_out/bench_c2h.log; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_c2h.log )
( timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3h.json 2> gpurun_out/bench_c3h.log; echo "bench c3 rc=$?"; grep -v "warmup" gpurun_out/bench_c3h.log | tail -3 )
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2h.json", "gpurun_out/bench_c3h.json"):
  try:
    d = json.load(open(f))
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2),
          "enc", {k: round(v) for k, v in d["extra"]["encode"].items() if k.endswith("MB_per_s")}, "parity", d["parity"]["equal"], "retried", len(d["retried_steps"]))
    rs = d["roofline"]["resident_split"]
    print("   ", rs["local_by_log_entries(<=512,<=4096,<=32768,more)"], rs["local_us_per_merge"], rs["grid_us_per_merge"], d.get("cpu_baseline", {}).get("ours_on_same_sample"))
  except Exception as e:
    print(f, "unreadable:", e)
PY
timeout 600 python scripts/profile_r2.py > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.log && cat gpurun_out/profile_plain.json &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:merge_cluster' -c 1 --csv --log-file gpurun_out/r2_merge_cluster_dram.csv python scripts/profile_r2.py > gpurun_out/ncu_m1.log 2>&1
echo "ncu dram rc=$?"; tail -2 gpurun_out/r2_merge_cluster_dram.csv
timeout 900 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none -k 'regex:merge_cluster' -c 1 --csv --log-file gpurun_out/r2_merge_cluster_sm.csv python scripts/profile_r2.py > gpurun_out/ncu_m2.log 2>&1
echo "ncu sm rc=$?"; tail -2 gpurun_out/r2_merge_cluster_sm.csv
