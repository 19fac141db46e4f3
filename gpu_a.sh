cd $GRAFT_REPO_ROOT
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
for nc in 0 1; do
SWB_NO_CLUSTER=$nc timeout 60 python scripts/profile_step.py config2_1GB 3 2>/dev/null | python -c "
import json,sys
best=None
for ln in sys.stdin:
  d=json.loads(ln); s=d['stats']
  if best is None or d['merge']<best[0]: best=(d['merge'], d, s)
if best:
  _,d,s=best
  print('no_cluster=$nc', 'best merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in ('host_pop_ms','host_wait_ms','host_apply_ms')})
else: print('no_cluster=$nc FAILED')
"
done
