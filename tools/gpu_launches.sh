#!/bin/bash
# ncu launch lists (device time per launch; cold-cache, serialised: compare SHARES) of one short bench run per mode
mkdir -p gpurun_out
for MODE in infer train; do
  BCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-graph --mode $MODE"
  $BCMD > gpurun_out/plain_$MODE.log 2>&1 || { echo "plain $MODE failed"; tail -5 gpurun_out/plain_$MODE.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$MODE.csv $BCMD > gpurun_out/ncu_launch_$MODE.log 2>&1
  tail -1 gpurun_out/ncu_launch_$MODE.log | cut -c1-160
  wc -l gpurun_out/launches_$MODE.csv
done
