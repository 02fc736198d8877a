#!/bin/bash
# DRAM traffic of every tensor-core conv launch of one inference step (for roofline.traffic)
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --mode infer"
$BCMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv3x3_tc|fcomb_tc" --csv --log-file gpurun_out/traffic.csv $BCMD > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-200
wc -l gpurun_out/traffic.csv
