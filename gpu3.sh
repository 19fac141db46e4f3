cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_train_parity.py -x -q -m gpu 2>&1 | tail -2
SWB_TRACE_WAIT=1 python scripts/profile_step.py config2_1GB 1 2>&1 | grep -E "trace" | tail -12
