cd $GRAFT_REPO_ROOT
N=${1:-2}
if [ "$N" = "2" ]; then ( timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -3 ); fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log; echo "bench$N rc=$?"
grep "bound" gpurun_out/bench_n$N.log | head -3
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['extra']['phase_ms'], d['extra']['e2e_phase_ms'], d['extra']['us_per_merge'], d['extra']['collectives_per_step'], d['extra']['unique_words'], d['clocks'])"
