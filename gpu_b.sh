cd $GRAFT_REPO_ROOT
SWB_KERNEL_TRACE=1 python -c "
from shredword_b200 import build as B; B.build(force=True)"
SWB_NO_PERSISTENT=1 SWB_TRACE_WAIT=1 python scripts/profile_step.py config2_1GB 2 2>&1 | grep -v '^{' | tail -30 | cut -c1-600
