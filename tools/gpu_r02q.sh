#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_step_differential.py tests/test_gpu_train.py tests/test_gpu_architectures.py tests/test_gpu_steps.py -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | grep -v "^frame\|^$" | tail -8
