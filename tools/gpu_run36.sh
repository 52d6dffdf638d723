cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l)
T=r3w
timeout 900 python -m pytest tests -m gpu -x -q -k "strip" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $NG --steps 5 2> gpurun_out/${T}_bench_${NG}gpu.err | grep '^{' > gpurun_out/${T}_bench_${NG}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_${NG}gpu.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value'])); print({k:(round(v['ms_per_panorama'],3), v['all_ranks_match_undivided']) for k,v in d['strip_split']['modes'].items()}, d['strip_split']['single_gpu_ms_per_panorama'])"
