#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_architectures.py tests/test_gpu_guards.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -v "^$" | tail -60
