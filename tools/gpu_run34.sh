cd $GRAFT_REPO_ROOT
T=${1:-r3s}
for V in 8 7 8 7; do
  export PANO_RESIZE_OCC=$V
  python bench.py --no-cpu-baseline --steps 5 --no-also --e2e-steps 2 2>/dev/null | grep '^{' > gpurun_out/${T}_bench_$V.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_$V.json'))
k = d['roofline']['kernels']
print('resize occ $V', 'value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), {n: round(v['ms_per_launch'], 3) for n, v in k.items() if n.startswith('fe_')})
PY
done
