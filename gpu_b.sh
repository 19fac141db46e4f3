cd $GRAFT_REPO_ROOT
for nc in 0 1; do
SWB_NO_CLUSTER=$nc SWB_TRACE_WAIT=1 timeout 60 python scripts/profile_step.py config2_1GB 2 > gpurun_out/trace$nc.log 2>&1
grep "trace\] cluster\|trace\] resident" gpurun_out/trace$nc.log | tail -1 | cut -c1-330
tail -1 gpurun_out/trace$nc.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('no_cluster=$nc', 'merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in ('host_pop_ms','host_wait_ms','host_apply_ms')})
"
done
