cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_dist.py -x -q -m gpu -k "range_split or (True and ascii)" 2>&1 | tail -15
