cd $GRAFT_REPO_ROOT
T=${1:-r3c}
CMD="python bench.py --workload config3 --steps 1 --warmup 3 --waves-per-step 1 --no-cpu-baseline --no-also --e2e-steps 1"
$CMD > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"blend_g0|warp_tile" -s 6 -c 2 -o gpurun_out/${T}_c3 $CMD > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_c3.ncu-rep
