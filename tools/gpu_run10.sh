cd $GRAFT_REPO_ROOT
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/strip_split_nccl.py --steps 30 > gpurun_out/r2j_strip_${N}gpu.json 2> gpurun_out/r2j_strip_${N}gpu.err
tail -2 gpurun_out/r2j_strip_${N}gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_strip_${N}gpu.json')); print(d['n_gpus'], 'single', round(d['single_gpu_ms_per_panorama'],3), {k:(round(v['ms_per_panorama'],3), v['all_ranks_match_undivided']) for k,v in d['modes'].items()})"
