cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/profile_step.py config2_1GB 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stats']
print({k:round(v,4) for k,v in d.items() if k!='stats'})
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in s.items()})
"
