cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -15
