# round 2, GPU call P (4 GPUs): the driver's own N=4 and N=2 commands (strong scaling, config 3, 20 + 5 steps)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for N in 4 2; do
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}p.json 2> gpurun_out/bench_n${N}p.log; echo "bench n$N rc=$?"; grep -v "warmup" gpurun_out/bench_n${N}p.log | tail -3 | cut -c1-300 )
python - <<PY
import json
try:
  d = json.load(open("gpurun_out/bench_n${N}p.json"))
  print("N=$N value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 3), d["extra"]["phase_ms"], d["extra"]["e2e_phase_ms"], "enc", {k: round(v) for k, v in d["extra"]["encode"].items() if k.endswith("MB_per_s")}, "parity", d["parity"]["equal"], d["scaling"], d["clocks"], "retried", d["retried_steps"])
except Exception as e:
  print("bench_n$N unreadable", e)
PY
done
