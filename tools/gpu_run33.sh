cd $GRAFT_REPO_ROOT
T=${1:-r3p}
for V in 3 2; do
  export PANO_C8_OCC=$V
  python bench.py --no-cpu-baseline --steps 5 --no-also --e2e-steps 2 2>/dev/null | grep '^{' > gpurun_out/${T}_bench_$V.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_$V.json'))
k = d['roofline']['kernels']
print('c8 occ $V', 'value', round(d['value']), 'ms/wave', round(d['ms_per_step']/16, 3), {n: round(v['ms_per_launch'], 3) for n, v in k.items() if n.startswith('collapse')})
PY
done
