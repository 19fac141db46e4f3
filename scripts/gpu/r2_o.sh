# round 2, GPU call O (1 GPU): pre-passes with the cheap-test / general-rule split
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_pretok.py tests/test_normalize.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/pytest_o.log; tail -4 gpurun_out/pytest_o.log )
( timeout 300 python scripts/profile_prepass.py 1000000000 > gpurun_out/prepass_o.json 2> gpurun_out/prepass_o.log; cat gpurun_out/prepass_o.json )
( timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:stream_map' -c 6 -o gpurun_out/r2_prepass python scripts/profile_prepass.py > gpurun_out/ncu_prepass.log 2>&1; echo "ncu prepass rc=$?"; tail -2 gpurun_out/ncu_prepass.log )
