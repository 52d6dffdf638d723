cd $GRAFT_REPO_ROOT
T=${1:-r2t}
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pageable or golden" 2>&1 | tail -5) > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
: > gpurun_out/${T}_pageable.jsonl
run() { env "$@" python tools/experiments/pageable_latency.py 2>/dev/null | grep '^{' >> gpurun_out/${T}_pageable.jsonl; }
run PANO_X=default
run PANO_HOST_NO_STREAM=1
run PANO_HOST_THREADS=2
run PANO_HOST_THREADS=3
run PANO_HOST_THREADS=6
run PANO_HOST_THREADS=8
run PANO_NO_HOST_STAGING=1
cat gpurun_out/${T}_pageable.jsonl
