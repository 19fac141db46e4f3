# round 2, GPU call F (8 GPUs): the driver's own N=8 command on the new bench (strong scaling, config 3), then config 5 at full size
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.log; echo "bench n8 rc=$?"; grep -v "warmup" gpurun_out/bench_n8.log | tail -12 )
python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/bench_n8.json"))
  print("N=8 value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], d["extra"]["e2e_phase_ms"], "enc", d["extra"].get("encode"), "parity", d["parity"]["equal"], d["scaling"], d["clocks"])
except Exception as e:
  print("bench_n8 unreadable", e)
PY
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 scripts/config5_check.py config5_50GB > gpurun_out/config5_n8.json 2> gpurun_out/config5_n8.log; echo "config5 rc=$?"; tail -5 gpurun_out/config5_n8.log; cat gpurun_out/config5_n8.json | cut -c1-1500 )
