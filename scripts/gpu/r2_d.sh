# round 2, GPU call D (1 GPU): encoder two-phase + normalize tests, bench config 2, ncu of the scripted merge_cluster launch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_encode_parity.py tests/test_normalize.py tests/test_gpu_fuzz.py -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_d.log; tail -6 gpurun_out/pytest_d.log )
( timeout 900 python bench.py --workload config2_1GB --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2d.json 2> gpurun_out/bench_c2d.log; echo "bench c2 rc=$?"; tail -2 gpurun_out/bench_c2d.log )
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2d.json"))
print("value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], "enc", d["extra"].get("encode"))
PY
timeout 600 python scripts/profile_r2.py > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.log && cat gpurun_out/profile_plain.json &&
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:merge_cluster|enc_fused' -c 2 -o gpurun_out/r2_merge_cluster python scripts/profile_r2.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
