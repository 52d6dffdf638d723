cd $GRAFT_REPO_ROOT
T=${1:-r2v}
(timeout 900 python -m pytest tests -m gpu -x -q -k "front or config2 or frontend or nvcam or pageable" 2>&1 | tail -5) > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
for V in 4 2 1; do
  export PANO_RESIZE_COLS=$V
  if [ $V = 4 ]; then (timeout 900 python -m pytest tests -m gpu -x -q -k "front or config2 or frontend or nvcam" 2>&1 | tail -2); fi
  timeout 900 python bench.py --no-cpu-baseline --steps 5 --no-also --e2e-steps 2 2> gpurun_out/${T}_bench_$V.err | grep '^{' > gpurun_out/${T}_bench_$V.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_$V.json'))
k = d['roofline']['kernels']
print('cols $V', 'value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), {n: round(v['ms_per_launch'], 3) for n, v in k.items() if n.startswith('fe_')})
PY
done
