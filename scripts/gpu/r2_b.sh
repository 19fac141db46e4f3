# round 2, GPU call B (1 GPU): forced-overflow variants + scripted launch after the spill fix, config 3 repeated, then the same under the debug-bounds build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_scale_parity.py::test_scripted_launch_equals_host_driven_loop -q -p no:cacheprovider -x 2>&1 | tail -80 > gpurun_out/pytest_b.log; tail -25 gpurun_out/pytest_b.log )
( timeout 900 python bench.py --steps 8 --warmup 3 --no-encode --no-cpu-baseline > gpurun_out/bench_c3b.json 2> gpurun_out/bench_c3b.log; echo "bench c3 rc=$?"; grep -v "^\[bench\] rank 0 warmup" gpurun_out/bench_c3b.log | tail -4 )
echo "== debug-bounds build"
SWB_DEBUG_BOUNDS=1 python shredword_b200/build.py --force > gpurun_out/build_dbg.log 2>&1; echo "build rc=$?"
( timeout 900 python -m pytest tests/test_gpu_variants.py -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_b_dbg.log; tail -12 gpurun_out/pytest_b_dbg.log )
( timeout 900 python bench.py --steps 6 --warmup 2 --no-encode --no-cpu-baseline > gpurun_out/bench_c3b_dbg.json 2> gpurun_out/bench_c3b_dbg.log; echo "bench c3 dbg rc=$?"; grep -v "^\[bench\] rank 0 warmup" gpurun_out/bench_c3b_dbg.log | tail -4 )
python - <<'PY'
import json
for f in ("gpurun_out/bench_c3b.json", "gpurun_out/bench_c3b_dbg.json"):
  try:
    d = json.load(open(f))
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "us/merge", round(d["extra"]["us_per_merge"], 2),
          "parity", d["parity"], d["roofline"]["resident_split"])
  except Exception as e:
    print(f, "unreadable:", e)
PY
