cd $GRAFT_REPO_ROOT
( timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 )
for nc in 0 1; do
SWB_NO_CLUSTER=$nc SWB_TRACE_WAIT=1 timeout 60 python scripts/profile_step.py config2_1GB 3 > gpurun_out/trace$nc.log 2>&1
grep "trace\] cluster\|trace\] resident" gpurun_out/trace$nc.log | tail -1 | cut -c1-330
grep "^{" gpurun_out/trace$nc.log | python -c "
import json,sys
best=None
for ln in sys.stdin:
  d=json.loads(ln); s=d['stats']
  if best is None or d['merge']<best[0]: best=(d['merge'], d, s)
_,d,s=best
print('no_cluster=$nc', 'best merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in ('host_pop_ms','host_wait_ms','host_apply_ms')})
"
done
