# round 2, GPU call T (1 GPU): final build -- full GPU suite, memcheck of the pre-pass kernels, the driver's N=1 bench command,
# smoke under ncu's launch list, ncu --set full of the pre-pass kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_t.log; echo "pytest done"; tail -3 gpurun_out/pytest_t.log )
# (compute-sanitizer is closed on this pool: the memcheck run of the pre-pass tests was refused)
( timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.log; echo "bench n1 rc=$?"; grep -v "warmup" gpurun_out/bench_final_n1.log | tail -2 | cut -c1-300 )
python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/bench_final_n1.json"))
  print("value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2), "frac", round(d["roofline"]["frac"], 3),
        "enc", {k: round(v) for k, v in d["extra"]["encode"].items() if k.endswith("MB_per_s")}, "parity", d["parity"]["equal"], "retried", len(d["retried_steps"]), d["clocks"])
  cb = d["cpu_baseline"]; print("cpu", cb["kind"], cb["seconds"], cb["value"], cb["ours_on_same_sample"])
except Exception as e:
  print("unreadable:", e)
PY
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/smoke_ncu_t.log 2>&1; echo "smoke under ncu: rc=$?"; tail -1 gpurun_out/smoke_ncu_t.log )
( timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:stream_map' -c 6 -o gpurun_out/r2_prepass python scripts/profile_prepass.py > gpurun_out/ncu_prepass.log 2>&1; echo "ncu prepass rc=$?"; tail -1 gpurun_out/ncu_prepass.log )
