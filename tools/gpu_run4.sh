cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
