cd $GRAFT_REPO_ROOT
SWB_TRACE_INIT=1 timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -5
( time timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) 2>&1
for nc in 0 1; do
SWB_NO_CLUSTER=$nc SWB_TRACE_WAIT=1 timeout 120 python scripts/profile_step.py config2_1GB 2 > gpurun_out/trace$nc.log 2>&1
grep "trace\|rror" gpurun_out/trace$nc.log | tail -3 | cut -c1-400
tail -1 gpurun_out/trace$nc.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('no_cluster=$nc', 'load', round(d['load'],3), 'merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in ('host_pop_ms','host_launch_ms','host_wait_ms','host_apply_ms')}, s['records'], s['kernel_launches'])
"
done
