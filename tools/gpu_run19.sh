cd $GRAFT_REPO_ROOT
T=${1:-r2s}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
PANO_DEBUG=1 python bench.py --no-cpu-baseline --steps 3 --no-also --e2e-steps 2 2>&1 >/dev/null | grep "panob200\] level" | sort | uniq -c > gpurun_out/${T}_tiles.txt
cat gpurun_out/${T}_tiles.txt
python bench.py 2> gpurun_out/${T}_bench.err | grep '^{' > gpurun_out/${T}_bench.json
PANO_NO_HOST_STAGING=1 python bench.py --no-cpu-baseline --steps 3 --e2e-steps 2 2> gpurun_out/${T}_bench_nostage.err | grep '^{' > gpurun_out/${T}_bench_nostage.json
for t in 1 2 8; do PANO_HOST_THREADS=$t python bench.py --no-cpu-baseline --steps 3 --e2e-steps 2 2>/dev/null | grep '^{' > gpurun_out/${T}_bench_threads$t.json; done
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench.json'))
print('value', round(d['value']), 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'parity', d.get('parity'))
print('latency', d['latency'])
print('cpu', json.dumps(d.get('cpu_baseline'))[:600])
print('also', json.dumps(d.get('also'))[:1800])
for n in ('nostage', 'threads1', 'threads2', 'threads8'):
    e = json.load(open('gpurun_out/${T}_bench_%s.json' % n))
    print(n, e['latency']['pageable_host_buffers'], 'config1', e['also']['config1']['latency']['pageable_host_buffers'])
PY
