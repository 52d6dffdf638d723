cd $GRAFT_REPO_ROOT
T=${1:-r2x}
for V in 5 6 7; do
  export PANO_RESIZE_OCC=$V
  timeout 900 python bench.py --no-cpu-baseline --steps 5 --no-also --e2e-steps 2 2> gpurun_out/${T}_bench_$V.err | grep '^{' > gpurun_out/${T}_bench_$V.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_$V.json'))
k = d['roofline']['kernels']
print('occ $V', 'value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), {n: round(v['ms_per_launch'], 3) for n, v in k.items() if n.startswith('fe_')})
PY
done
