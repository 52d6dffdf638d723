cd $GRAFT_REPO_ROOT
T=${1:-r2z}
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
: > gpurun_out/${T}_latency.jsonl
run() { env "$@" python tools/experiments/pageable_latency.py 2>/dev/null | grep '^{' >> gpurun_out/${T}_latency.jsonl; }
run PANO_X=default
run PANO_NO_OVERLAP=1
run PANO_NO_GRAPH=1
run PANO_HOST_STREAM_OUT=1
cat gpurun_out/${T}_latency.jsonl
