cd $GRAFT_REPO_ROOT
T=${1:-r3i}
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
python bench.py 2> gpurun_out/${T}_bench.err | grep '^{' > gpurun_out/${T}_bench.json
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench.json'))
print('value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), 'e2e', round(d['e2e']['value']), 'parity', d['parity']['max_abs_diff'] if d.get('parity') else None)
print({n: round(v['ms_per_launch'], 3) for n, v in d['roofline']['kernels'].items()})
l = d['latency']; print('latency', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in l.items() if 'ms' in k}, l['pageable_host_buffers']['pano_process_ms_p50'])
a = d['also']
print('config1', round(a['config1']['value']), round(a['config1']['ms_per_wave'], 3), a['config1']['latency']['device_ms_per_frame_set'])
print('config3', a.get('config3'))
print('config5', round(a['config5']['value']), round(a['config5']['config1_tables']['value']))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('cached_maps', {}).get('value'))
PY
tail -3 gpurun_out/${T}_bench.err
