#!/bin/bash
# quick GPU visit: parity tests + bench line (no ncu)
mkdir -p gpurun_out
for f in tests/test_gpu_conv.py tests/test_gpu_punet.py; do
  echo "=== $f"
  timeout 600 python -m pytest $f -m gpu -q --timeout 180 -p no:cacheprovider -s 2>&1 | tail -${TAILN:-30}
done
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} 2>&1 | tail -3 | tee gpurun_out/bench.json
