# round 2, GPU call C (1 GPU): full -m gpu suite with the new tokenizer / single-pass encoder, bench config 2, ncu captures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log )
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; echo "pytest done"; tail -8 gpurun_out/pytest_gpu.log )
( timeout 900 python bench.py --workload config2_1GB --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.log; echo "bench c2 rc=$?"; tail -2 gpurun_out/bench_c2.log )
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2.json",):
  try:
    d = json.load(open(f))
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2),
          "tok", d["extra"]["tokenize"], "enc", d["extra"].get("encode"), "parity", d["parity"]["equal"], d["roofline"]["frac"], d["roofline"]["resident_split"])
  except Exception as e:
    print(f, "unreadable:", e)
PY
timeout 600 python scripts/profile_r2.py > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.log && cat gpurun_out/profile_plain.json &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:wt_|count_rows|merge_cluster|enc_|tokfreq|ip_index|pt_emit' -c 200 --csv --log-file gpurun_out/r2_launches_profile.csv python scripts/profile_r2.py > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:merge_cluster|wt_tokenize|enc_fused' -c 3 -o gpurun_out/r2_hot_kernels python scripts/profile_r2.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -12
