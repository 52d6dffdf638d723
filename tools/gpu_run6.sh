cd $GRAFT_REPO_ROOT
B="python bench.py --no-also --no-cpu-baseline --steps 5 --e2e-steps 2"
PANO_WARP_OCC=4 $B > gpurun_out/r2f_w4.json 2>/dev/null
PANO_WARP_OCC=6 $B > gpurun_out/r2f_w6.json 2>/dev/null
PANO_WARP_OCC=4 $B --workload config1 > gpurun_out/r2f_c1w4.json 2>/dev/null
PANO_WARP_OCC=6 $B --workload config1 > gpurun_out/r2f_c1w6.json 2>/dev/null
PANO_DOWN_BAND=32 $B > gpurun_out/r2f_band32.json 2>/dev/null
PANO_DOWN_BAND=16 PANO_DOWN_OCC=6 $B > gpurun_out/r2f_band16o6.json 2>/dev/null
PANO_DOWN_BAND=24 $B > gpurun_out/r2f_band24.json 2>/dev/null
python - <<'PY'
import json
for n in ['w4','w6','c1w4','c1w6','band32','band16o6','band24']:
    try:
        d=json.load(open('gpurun_out/r2f_%s.json'%n)); k=d['roofline']['kernels']
        print(n, round(d['value']), ' '.join('%s=%.3f'%(a,k[a]['ms_per_launch']) for a in ['warp','pyrdown_l0','pyrdown_l1','pyrdown_l2','collapse_l0','collapse_l1'] if a in k))
    except Exception as e: print(n,'ERR',e)
PY
