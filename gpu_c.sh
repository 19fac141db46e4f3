cd $GRAFT_REPO_ROOT
timeout 120 python scripts/profile_step.py 250000000 2 2>&1 | tail -2 | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_train_parity.py -x -q -k "ref_fixture or unk_enters or ragged or self_pairs" > gpurun_out/sanitizer.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|at 0x|merge_cluster|=========     at" gpurun_out/sanitizer.log | head -20
