cd $GRAFT_REPO_ROOT
T=${1:-r2p}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
for V in pdl nopdl; do
  if [ $V = nopdl ]; then export PANO_NO_PDL=1; fi
  timeout 900 python bench.py --no-cpu-baseline 2> gpurun_out/${T}_bench_$V.err | grep '^{' > gpurun_out/${T}_bench_$V.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_$V.json'))
print('$V', 'value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), 'e2e', round(d['e2e']['value']))
l = d['latency']; print(' latency', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in l.items() if 'ms' in k})
a = d['also']; print(' config1', round(a['config1']['value']), {k: round(v, 4) for k, v in a['config1']['latency'].items() if 'ms' in k and isinstance(v, float)})
print(' config5', round(a['config5']['value']), round(a['config5']['host_streamed']['value']))
PY
done
