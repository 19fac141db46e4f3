cd $GRAFT_REPO_ROOT
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -1
for nb in 1 0; do
SWB_NO_BIRTH_LOG=$nb SWB_TRACE_WAIT=1 python scripts/profile_step.py config2_1GB 2 > gpurun_out/trace$nb.log 2>&1
grep "trace" gpurun_out/trace$nb.log | tail -2 | cut -c1-400
tail -1 gpurun_out/trace$nb.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('no_birth_log=$nb', 'load', round(d['load'],3), 'merge', round(d['merge'],4), 'us/merge', round(d['us_per_merge'],2), {k:round(s[k],1) for k in ('host_pop_ms','host_launch_ms','host_wait_ms','host_apply_ms')}, s['records'], s['kernel_launches'])
"
done
