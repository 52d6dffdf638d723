cd $GRAFT_REPO_ROOT
T=${1:-r3v}
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
python tools/experiments/pageable_latency.py 2>/dev/null | grep "^{" | tee gpurun_out/${T}_latency.jsonl
PANO_NO_FORK=1 python tools/experiments/pageable_latency.py 2>/dev/null | grep "^{" | tee -a gpurun_out/${T}_latency.jsonl
