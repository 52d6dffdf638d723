cd $GRAFT_REPO_ROOT
T=${1:-r3a}
(timeout 900 python -m pytest tests -m gpu -x -q -k "feather or no_blend or config3 or gain or cxx" 2>&1 | tail -5) > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --workload config3 --no-cpu-baseline --steps 5 --no-also --e2e-steps 2 2> gpurun_out/${T}_bench_c3.err | grep '^{' > gpurun_out/${T}_bench_c3.json
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_c3.json'))
k = d['roofline']['kernels']
print('config3 value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), {n: round(v['ms_per_launch'], 3) for n, v in k.items()})
PY
