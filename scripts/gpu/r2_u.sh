# round 2, GPU call U (1 GPU): tokeniser flag check only on long probe sequences -- load times (10 GB, 32 MB sample, forced growth)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_variants.py -q -m gpu -p no:cacheprovider -k "word_table_growth or default" 2>&1 | tail -2 )
( timeout 300 python scripts/small_load_probe.py 2>&1 | head -3 )
( timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-encode > gpurun_out/bench_c3u.json 2> gpurun_out/bench_c3u.log; echo "bench c3 rc=$?" )
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c3u.json")); e = d["extra"]
print("value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "tokenize", e["tokenize"]["ms"], e["phase_ms"], "parity", d["parity"]["equal"])
PY
