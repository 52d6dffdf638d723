cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
B="python bench.py --no-also --no-cpu-baseline --steps 5 --e2e-steps 2"
$B --workload config3 > gpurun_out/r2m_c3.json 2>/dev/null
PANO_FEATHER_FLOAT=1 $B --workload config3 > gpurun_out/r2m_c3_float.json 2>/dev/null
python - <<'PY'
import json
for n in ['c3','c3_float']:
    try:
        d=json.load(open('gpurun_out/r2m_%s.json'%n)); k=d['roofline']['kernels']
        print(n, round(d['value']), ' '.join('%s=%.3f(%.2f)'%(a,v['ms_per_launch'],v['frac']) for a,v in k.items()))
    except Exception as e: print(n,'ERR',e)
PY
