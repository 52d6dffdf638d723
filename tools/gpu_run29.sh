cd $GRAFT_REPO_ROOT
T=${1:-r3j}
for rep in 1 2; do
for V in cur prev; do
  if [ $V = prev ]; then export PANOB200_LIB=$GRAFT_REPO_ROOT/img-stitching_b200/lib/libpanob200_prev.so; else unset PANOB200_LIB; fi
  python bench.py --no-cpu-baseline --no-also --steps 5 2>/dev/null | grep '^{' > gpurun_out/${T}_$V$rep.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${T}_$V$rep.json')); e = d['e2e']
print('$V$rep', 'value', round(d['value']), 'e2e', round(e['value']), 'ceiling', round(e['pcie_ceiling_panoramas_per_s']), 'ratio', round(e['value']/e['pcie_ceiling_panoramas_per_s'], 3))
PY
done; done
