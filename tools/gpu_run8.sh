cd $GRAFT_REPO_ROOT
for N in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/experiments/pcie_ranks.py > gpurun_out/r2h_pcie_${N}.json 2> gpurun_out/r2h_pcie_${N}.err
  tail -c 700 gpurun_out/r2h_pcie_${N}.json; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err
tail -3 gpurun_out/r2h_bench_8gpu.err; head -c 400 gpurun_out/r2h_bench_8gpu.json
