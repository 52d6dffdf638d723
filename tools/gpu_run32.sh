cd $GRAFT_REPO_ROOT
T=${1:-r3o}
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
PANO_DEBUG=1 python bench.py --no-cpu-baseline --steps 5 --no-also --e2e-steps 2 2> gpurun_out/${T}_bench.err | grep '^{' > gpurun_out/${T}_bench.json
grep "panob200\] level" gpurun_out/${T}_bench.err | sort | uniq -c
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench.json'))
k = d['roofline']['kernels']
print('value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), {n: round(v['ms_per_launch'], 3) for n, v in k.items() if n.startswith('collapse')})
PY
