cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l)
T=r2u
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tools/experiments/pcie_ranks.py 2> gpurun_out/${T}_pcie_${NG}.err | grep '^{' > gpurun_out/${T}_pcie_${NG}.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_pcie_${NG}.json')); print(d['n_gpus'], {k:round(v['aggregate_GBps'],1) for k,v in d.items() if isinstance(v,dict)})"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $NG --steps 10 2> gpurun_out/${T}_bench_${NG}gpu.err | grep '^{' > gpurun_out/${T}_bench_${NG}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_${NG}gpu.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['pcie_ceiling_GBps']); print({k:(round(v['ms_per_panorama'],3), v['all_ranks_match_undivided'], v.get('cameras_held_per_rank_max')) for k,v in d['strip_split']['modes'].items()}, d['strip_split']['single_gpu_ms_per_panorama']); print('config5', d['also']['config5'])"
tail -3 gpurun_out/${T}_bench_${NG}gpu.err
