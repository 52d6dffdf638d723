cd $GRAFT_REPO_ROOT
T=${1:-r2q}
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strip_rank_reads or strip_split_hybrid" 2>&1 | tail -5) > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
for V in 0 1184 8192 100000000; do
  export PANO_PDL_MAX_BLOCKS=$V
  timeout 900 python bench.py --no-cpu-baseline --steps 5 2> gpurun_out/${T}_bench_$V.err | grep '^{' > gpurun_out/${T}_bench_$V.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_$V.json'))
print('$V', 'value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), 'e2e', round(d['e2e']['value']))
l = d['latency']; print(' latency', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in l.items() if 'ms' in k})
a = d['also']; print(' config1', round(a['config1']['value']), {k: round(v, 4) for k, v in a['config1']['latency'].items() if 'ms' in k and isinstance(v, float)})
PY
done
