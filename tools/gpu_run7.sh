cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
B="python bench.py --no-also --no-cpu-baseline --steps 5 --e2e-steps 2"
$B --workload config1 > gpurun_out/r2g_c1.json 2>/dev/null
PANO_NO_TMA_BGR=1 $B --workload config1 > gpurun_out/r2g_c1_nobgrtma.json 2>/dev/null
$B > gpurun_out/r2g_c2.json 2>/dev/null
$B --workload config3 > gpurun_out/r2g_c3.json 2>/dev/null
python - <<'PY'
import json
for n in ['c1','c1_nobgrtma','c2','c3']:
    try:
        d=json.load(open('gpurun_out/r2g_%s.json'%n)); k=d['roofline']['kernels']
        print(n, round(d['value']), ' '.join('%s=%.3f'%(a,v['ms_per_launch']) for a,v in k.items() if v['ms_per_launch']>0.1))
    except Exception as e: print(n,'ERR',e)
PY
