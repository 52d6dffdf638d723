#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU):  bash tools/profile.sh <tag>
#  1. plain bench run (must exit 0), 2. ncu launch list of the same command,
#  3. ncu --set full of the hot kernels.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --waves-per-step 1 --no-cpu-baseline --no-also --e2e-steps 1"   # default shape: one wave of 64 frame-sets
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on \
    -k regex:"collapse8|collapse_walk|warp_tile|pyrdown8|cubic5|resize4_walk" -s 34 -c 17 -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
