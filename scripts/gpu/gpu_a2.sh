cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log; echo "bench2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['extra']['phase_ms'], d['extra']['e2e_phase_ms'], d['extra']['us_per_merge'])"
