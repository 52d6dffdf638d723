cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strip" 2>&1 | tail -30) > gpurun_out/r2i_pytest_strip.log
tail -6 gpurun_out/r2i_pytest_strip.log
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/strip_split_nccl.py --steps 20 > gpurun_out/r2i_strip_${N}gpu.json 2> gpurun_out/r2i_strip_${N}gpu.err
tail -2 gpurun_out/r2i_strip_${N}gpu.err; cat gpurun_out/r2i_strip_${N}gpu.json
