cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l)
T=r3n
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $NG --steps 10 2> gpurun_out/${T}_bench_${NG}gpu.err | grep '^{' > gpurun_out/${T}_bench_${NG}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_${NG}gpu.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['pcie_ceiling_panoramas_per_s'])); print({k:(round(v['ms_per_panorama'],3), v['all_ranks_match_undivided'], v.get('cameras_held_per_rank_max')) for k,v in d['strip_split']['modes'].items()}, d['strip_split']['single_gpu_ms_per_panorama']); a=d['also']; print('config1', round(a['config1']['value']), 'config3', a['config3'].get('value'), a['config3'].get('error'), 'config5', round(a['config5']['value']), round(a['config5']['config1_tables']['value']))"
tail -2 gpurun_out/${T}_bench_${NG}gpu.err
