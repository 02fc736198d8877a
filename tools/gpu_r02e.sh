#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -x -k "cta_pair" 2>&1 | tail -5
timeout 600 python tools/conv_layer_bench.py 2>&1 | tee gpurun_out/conv_layer_bench.md
PDA_CONV_PAIR=1 timeout 300 python bench.py --mode infer --no-extras --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
