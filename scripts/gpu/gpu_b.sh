cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 )
timeout 300 python scripts/determinism_check.py config2_1GB 6 2>&1 | tail -6
SWB_TRACE_WAIT=1 timeout 120 python scripts/profile_step.py config2_1GB 2 > gpurun_out/trace0.log 2>&1
grep "trace\] hints" gpurun_out/trace0.log | tail -1 | cut -c1-600
tail -1 gpurun_out/trace0.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('load', round(d['load'],4), 'merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in s if k.startswith('host_') or k.startswith('resident') or k.startswith('hints')})
"
