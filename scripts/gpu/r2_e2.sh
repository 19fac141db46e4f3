# round 2, GPU call E (2 GPUs): the 2-GPU parity tests (rank-0 loop, replicated, sharded, Python-driven exchange, range-split load) and the strong-scaling bench at N=2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( timeout 1500 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_dist.log; tail -8 gpurun_out/pytest_dist.log )
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log; echo "bench n2 rc=$?"; grep -v "warmup" gpurun_out/bench_n2.log | tail -6 )
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2.json"))
print("N=2 value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], d["extra"]["e2e_phase_ms"], "enc", d["extra"].get("encode"), "parity", d["parity"]["equal"], d["scaling"])
PY
