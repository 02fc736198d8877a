#!/bin/bash
# one GPU visit: parity tests, smoke, bench line, ncu launch list + full capture of the conv kernel
mkdir -p gpurun_out
for f in tests/test_gpu_conv.py tests/test_gpu_punet.py; do
  echo "=== $f"
  timeout 600 python -m pytest $f -m gpu -q --timeout 180 -p no:cacheprovider -s 2>&1 | tail -25
done
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
BCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "=== ncu launches"
$BCMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
echo "=== ncu full (conv kernel)"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 4 -o gpurun_out/prof_conv $BCMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
