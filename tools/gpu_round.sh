#!/bin/bash
# one GPU visit: parity tests, smoke, bench line (optionally ncu with NCU=1)
mkdir -p gpurun_out
for f in tests/test_gpu_conv.py tests/test_gpu_punet.py tests/test_gpu_train.py tests/test_gpu_steps.py; do
  echo "=== $f"
  timeout 240 python -m pytest $f -m gpu -q --timeout 60 -p no:cacheprovider -s -x 2>&1 | tail -${TAIL:-30}
done
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"; timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
if [ -n "$NCU" ]; then
BCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "=== ncu launches"
$BCMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
fi
