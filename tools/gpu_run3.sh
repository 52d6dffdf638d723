cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err
tail -5 gpurun_out/r2c_bench_2gpu.err; head -c 300 gpurun_out/r2c_bench_2gpu.json
