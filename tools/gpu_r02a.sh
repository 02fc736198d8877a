#!/bin/bash
# round 2, GPU visit a: every -m gpu test (new parity / differential tests report their numbers), smoke, the fused Fcomb
# A/B (w3 in registers / shared memory), full bench
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for v in "" _w0 _w64; do
  echo "=== fcomb sweep lib$v"
  PDA_B200_LIB=$PWD/probabilistic_domain_adaptation_b200/libpda_b200$v.so timeout 300 python tools/fcomb_sweep.py 8 16 64 2>&1 | tail -4 | tee gpurun_out/fcomb_sweep$v.md
done
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log > gpurun_out/bench.json; head -c 600 gpurun_out/bench.json; echo
ls -la gpurun_out | tail -12
