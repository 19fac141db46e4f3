# round 2, GPU call X (1 GPU): look-ahead reuse with validated entries -- parity tests of the trainer, then A/B config 3 depth 6 vs 2 (5 steps each, alternating)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_fuzz.py tests/test_gpu_scale_parity.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3 )
for D in 6 2 6 2; do
  ( SWB_PEEK_DEPTH=$D timeout 900 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-encode > gpurun_out/bench_c3x_$D.json 2> gpurun_out/bench_c3x_$D.log; echo "bench c3 depth $D rc=$?"
    python - <<PY
import json
d = json.load(open("gpurun_out/bench_c3x_$D.json")); e = d["extra"]
print("  depth $D: ms", round(d["ms_per_step"], 1), "merge_ms", round(e["phase_ms"]["merge_ms"], 1), "dev us", round(d["roofline"]["avg_launch_us"], 2), e["look_ahead"], e["host_split_ms"], "parity", d["parity"]["equal"])
PY
  )
done
