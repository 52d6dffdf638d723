cd $GRAFT_REPO_ROOT
B="python bench.py --no-also --no-cpu-baseline --steps 5 --e2e-steps 2"
$B > gpurun_out/r2e_default.json 2> gpurun_out/r2e_default.err
PANO_DOWN_OCC=6 $B > gpurun_out/r2e_down6.json 2>/dev/null
PANO_DOWN_OCC=8 $B > gpurun_out/r2e_down8.json 2>/dev/null
PANO_WALK_OCC=8 $B > gpurun_out/r2e_walk8.json 2>/dev/null
PANO_WALK_OCC=10 $B > gpurun_out/r2e_walk10.json 2>/dev/null
PANO_DOWN_BAND=16 $B > gpurun_out/r2e_band16.json 2>/dev/null
$B --workload config1 > gpurun_out/r2e_c1.json 2>/dev/null
python - <<'PY'
import json
for n in ['default','down6','down8','walk8','walk10','band16','c1']:
    try:
        d=json.load(open('gpurun_out/r2e_%s.json'%n)); k=d['roofline']['kernels']
        print(n, round(d['value']), ' '.join('%s=%.3f'%(a,k[a]['ms_per_launch']) for a in ['warp','pyrdown_l0','pyrdown_l1','collapse_l0','collapse_l1','collapse_l2'] if a in k))
    except Exception as e: print(n,'ERR',e)
PY
