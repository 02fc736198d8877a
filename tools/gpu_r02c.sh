#!/bin/bash
# round 2, GPU visit c: all -m gpu tests (fp16 no-grad path) without the CTA-pair conv, then the CTA-pair tests in their
# own process (a protocol bug traps and poisons the context), smoke, bench A/B
mkdir -p gpurun_out
echo "=== pytest -m gpu (not cta_pair)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "not cta_pair" 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== pytest cta_pair"; timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -x -k "cta_pair" 2>&1 | tail -15 | tee gpurun_out/pytest_pair.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log > gpurun_out/bench.json; head -c 400 gpurun_out/bench.json; echo
echo "=== bench infer, pair mode"; PDA_CONV_PAIR=1 timeout 600 python bench.py --mode infer --no-extras --no-cpu-baseline > gpurun_out/bench_pair.log 2>&1; tail -1 gpurun_out/bench_pair.log > gpurun_out/bench_pair.json; head -c 400 gpurun_out/bench_pair.json; echo
ls -la gpurun_out | tail -12
