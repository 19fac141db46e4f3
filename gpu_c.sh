cd $GRAFT_REPO_ROOT
for mode in "SWB_X=1"; do
echo "== $mode"
env $mode timeout 300 python scripts/determinism_check.py config2_1GB 12 2>&1 | tail -12
done
