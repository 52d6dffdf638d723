cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
bash tools/profile.sh r2l > gpurun_out/r2l_profile.log 2>&1
tail -3 gpurun_out/r2l_profile.log
