# round 2, GPU call A (1 GPU): smoke, the whole -m gpu suite, memcheck of smoke, bench on configs 2 and 3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
( timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log )
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; echo "pytest done"; tail -15 gpurun_out/pytest_gpu.log )
( SWB_HOST_TIMEOUT_MS=600000 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|smoke ok" gpurun_out/memcheck_smoke.log | head -10 )
( timeout 900 python bench.py --workload config2_1GB --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.log; echo "bench c2 rc=$?"; tail -3 gpurun_out/bench_c2.log )
( timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "bench c3 rc=$?"; tail -3 gpurun_out/bench_c3.log )
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2.json", "gpurun_out/bench_c3.json"):
  try:
    d = json.load(open(f))
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2),
          "tok", d["extra"]["tokenize"], "enc", d["extra"].get("encode", {}).get("MB_per_s"), d["extra"].get("encode", {}).get("e2e_MB_per_s"), "parity", d["parity"]["equal"], d["roofline"]["frac"], d["roofline"]["resident_split"])
    print("  cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("ours_on_same_sample"))
  except Exception as e:
    print(f, "unreadable:", e)
PY
