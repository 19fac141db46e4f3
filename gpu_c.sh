cd $GRAFT_REPO_ROOT
for m in 0 2048 4096 8192 16384; do
SWB_IP_LOCAL_MAX=$m SWB_TRACE_WAIT=1 timeout 120 python scripts/profile_step.py config2_1GB 3 > gpurun_out/trace_ip$m.log 2>&1
tail -1 gpurun_out/trace_ip$m.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('ipmax=$m load', round(d['load'],4), 'merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in s if k.startswith('resident') or k.startswith('hints_t')})
"
done
