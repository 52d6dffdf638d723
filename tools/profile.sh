#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU, ONE profiler pass per call):
#   bash tools/profile.sh <tag> launches   plain bench run (must exit 0), then the ncu launch list of the same command
#   bash tools/profile.sh <tag> full       plain bench run (must exit 0), then ncu --set full of the hot kernels of one wave
# Outputs land in gpurun_out/.
set -u
TAG=${1:-r1}
MODE=${2:-launches}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --waves-per-step 1 --no-cpu-baseline --no-also --e2e-steps 1"   # default shape: one wave of 64 frame-sets
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
if [ "$MODE" = launches ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
else
  ncu --set full --clock-control none --import-source on \
      -k regex:"collapse8|collapse_walk|warp_tile|pyrdown8|cubic5|resize4_walk" -s 34 -c 17 -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
ls -la gpurun_out | tail -6
