# round 2, GPU call Z2 (2 GPUs): the N>1 path of the final build once more (short config 2 run), with extra.exchange_ms
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --workload config2_1GB --steps 2 --warmup 3 --no-encode > gpurun_out/bench_n2_c2_final.json 2> gpurun_out/bench_n2_c2_final.log; echo rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2_c2_final.json")); e = d["extra"]
print(d["n_gpus"], round(d["value"], 2), round(d["ms_per_step"], 1), e["phase_ms"], "exchange_ms", e["exchange_ms"], e["exchange_bytes_per_step"], d["parity"]["equal"], d["scaling"])
PY
