# round 2, GPU call J (8 GPUs): the driver's own N=8 command (strong scaling, config 3, 20 + 5 steps) on the build with the timer-skew fix
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8j.json 2> gpurun_out/bench_n8j.log; echo "bench n8 rc=$?"; grep -v "warmup" gpurun_out/bench_n8j.log | tail -12 )
python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/bench_n8j.json"))
  print("N=8 value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], d["extra"]["e2e_phase_ms"], "enc", d["extra"].get("encode"), "parity", d["parity"]["equal"], d["scaling"], d["clocks"], "retried", d["retried_steps"])
except Exception as e:
  print("bench_n8 unreadable", e)
PY
