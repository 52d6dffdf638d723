cd $GRAFT_REPO_ROOT
T=${1:-r2y}
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
