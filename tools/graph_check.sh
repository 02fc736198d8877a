#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_steps.py tests/test_gpu_eager_compare.py -m gpu -q --timeout 200 -p no:cacheprovider -x -s 2>&1 | tail -25
