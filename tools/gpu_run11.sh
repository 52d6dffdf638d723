cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
B="python bench.py --no-cpu-baseline --steps 5 --e2e-steps 2"
$B > gpurun_out/r2k_c2.json 2>/dev/null
PANO_NO_TOP_FUSED=1 $B --no-also > gpurun_out/r2k_c2_notop.json 2>/dev/null
python - <<'PY'
import json
for n in ['c2','c2_notop']:
    try:
        d=json.load(open('gpurun_out/r2k_%s.json'%n)); k=d['roofline']['kernels']
        print(n, round(d['value']), 'launches/wave', d['gpu_launches_per_wave'], ' '.join('%s=%.3f'%(a,v['ms_per_launch']) for a,v in k.items()))
        if d.get('latency'): print('  latency', d['latency']['pano_process_ms_p50'], d['latency']['device_ms_per_frame_set'], d['latency']['launches_per_call'])
        if d.get('also'): print('  c1', round(d['also']['config1']['value']), d['also']['config1']['latency']['device_ms_per_frame_set'])
    except Exception as e: print(n,'ERR',e)
PY
