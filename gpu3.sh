cd $GRAFT_REPO_ROOT
for mode in 0 1 0 1; do
SWB_HOST_TABLE=$mode python scripts/profile_step.py config2_1GB 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('host_table=$mode', 'load', round(d['load'],3), 'merge', round(d['merge'],4), {k:round(s[k],1) for k in ('host_pop_ms','host_launch_ms','host_wait_ms','host_apply_ms')}, s['records'])
"
done
./scripts/latency_probe | head -3
