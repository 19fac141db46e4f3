cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_train_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python scripts/determinism_check.py config2_1GB 2 2>&1 | tail -2
