cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_train_parity.py -x -q -m gpu -k "resident_loop or ascii_ties or pipelined" 2>&1 | tail -4
