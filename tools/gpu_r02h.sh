#!/bin/bash
mkdir -p gpurun_out
for WM in 1 2; do
  echo "=== PDA_CONV_WIDE=$WM pair tests"
  PDA_CONV_WIDE=$WM timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -k "cta_pair" 2>&1 | tail -4
done
for WM in 0 1; do
  echo "=== PDA_CONV_WIDE=$WM layer bench (pair kernel)"
  PDA_CONV_WIDE=$WM timeout 600 python tools/conv_layer_bench.py --pair 1 2>&1 | tee gpurun_out/conv_layer_bench_wide$WM.md
done
