#!/bin/bash
# ncu evidence for one build: launch list of a short bench run + full captures of selected kernels.
# usage: KERNELS="fcomb_tc conv3x3_tc" MODE=infer bash tools/gpu_ncu.sh
mkdir -p gpurun_out
MODE=${MODE:-infer}
BCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --mode $MODE"
$BCMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$MODE.csv $BCMD > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
for k in ${KERNELS:-fcomb_tc}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-6} -c ${COUNT:-2} -f -o gpurun_out/prof_$k $BCMD > gpurun_out/ncu_full_$k.log 2>&1
  tail -1 gpurun_out/ncu_full_$k.log | cut -c1-200
done
ls -la gpurun_out | tail -12
