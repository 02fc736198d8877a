#!/bin/bash
# full GPU visit: every -m gpu test, smoke, both bench arms, ncu launch lists (inference + training) and full captures
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -4
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log > gpurun_out/bench.json; head -c 400 gpurun_out/bench.json; echo
echo "=== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | head -c 600; echo
if [ -n "$NCU" ]; then
for MODE in infer train; do
  BCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-graph --mode $MODE"
  $BCMD > gpurun_out/plain_$MODE.log 2>&1 || { echo "plain $MODE failed"; tail -5 gpurun_out/plain_$MODE.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$MODE.csv $BCMD > gpurun_out/ncu_launch_$MODE.log 2>&1
  tail -1 gpurun_out/ncu_launch_$MODE.log | cut -c1-160
done
BCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-graph --mode infer"
for k in conv3x3_tc fcomb_tc; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 2 -f -o gpurun_out/prof_$k $BCMD > gpurun_out/ncu_full_$k.log 2>&1
  tail -1 gpurun_out/ncu_full_$k.log | cut -c1-160
done
fi
ls -la gpurun_out | tail -8
