#!/bin/bash
# runs each GPU test file in its own process so one sticky CUDA error cannot poison the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_conv.py tests/test_gpu_punet.py; do
  echo "=== $f"
  timeout 600 python -m pytest $f -m gpu -q --timeout 180 -p no:cacheprovider 2>&1 | tail -70
done
