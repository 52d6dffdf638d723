cd $GRAFT_REPO_ROOT
T=${1:-r2o}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
python bench.py --no-cpu-baseline 2> gpurun_out/${T}_bench.err | grep '^{' > gpurun_out/${T}_bench.json
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench.json'))
print('value', round(d['value']), 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'parity', d.get('parity'))
print(json.dumps(d.get('kernels') or d.get('breakdown') or {}, indent=0)[:3000])
print('also', json.dumps(d.get('also'))[:600])
PY
