cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_train_parity.py -x -q -m gpu 2>&1 | tail -4
for mode in 0 1; do
SWB_NO_PERSISTENT=$mode timeout 300 python scripts/profile_step.py config2_1GB 2 2>&1 | tail -1 | cut -c1-3000 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print('no_persistent=$mode', 'load', round(d['load'],3), 'merge', round(d['merge'],4), {k:round(s[k],1) for k in ('host_pop_ms','host_launch_ms','host_wait_ms','host_apply_ms')}, s['records'], s['kernel_launches'])
"
done
