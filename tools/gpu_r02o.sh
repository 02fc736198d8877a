#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_punet.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -v "^frame\|^$" | head -60
PDA_FUSE_UPSAMPLE=1 timeout 600 python bench.py --mode infer --no-extras --no-cpu-baseline 2>&1 | grep -v "^frame" | tail -25
