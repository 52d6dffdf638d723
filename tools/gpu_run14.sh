cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/experiments/pcie_ranks.py 2> gpurun_out/r2n_pcie_${N}.err | grep '^{' > gpurun_out/r2n_pcie_${N}.json
  python -c "
import json; d=json.load(open('gpurun_out/r2n_pcie_${N}.json')); print(d['n_gpus'], {k:round(v['aggregate_GBps'],1) for k,v in d.items() if isinstance(v,dict)})"
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 10 2> gpurun_out/r2n_bench_${NG}gpu.err | grep '^{' > gpurun_out/r2n_bench_${NG}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_${NG}gpu.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['pcie_ceiling_GBps']); print({k:(round(v['ms_per_panorama'],3), v['all_ranks_match_undivided']) for k,v in d['strip_split']['modes'].items()}, d['strip_split']['single_gpu_ms_per_panorama'])"
