cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -12
