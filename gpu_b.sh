cd $GRAFT_REPO_ROOT
for nh in 0; do
SWB_NO_HINTS=$nh SWB_TRACE_WAIT=1 timeout 120 python scripts/profile_step.py config2_1GB 3 > gpurun_out/trace$nh.log 2>&1
grep "trace\]" gpurun_out/trace$nh.log | tail -4 | cut -c1-400
tail -1 gpurun_out/trace$nh.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('no_hints=$nh merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in s if k.startswith('host_') or k.startswith('resident') or k.startswith('hints') or k.startswith('heap')})
"
done
