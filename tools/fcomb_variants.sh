#!/bin/bash
# fcomb kernel variants (PDA_FCOMB_VARIANT): parity tests + S sweep for each; ncu full capture of S=16 for NCU_VARS
mkdir -p gpurun_out
for v in ${VARS:-0 1 3}; do
  echo "=== variant $v"
  PDA_FCOMB_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_punet.py -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -4
  PDA_FCOMB_VARIANT=$v timeout 120 python tools/fcomb_sweep.py ${SWEEP:-1 4 16 64} 2>&1 | tail -6
done
for v in ${NCU_VARS:-}; do
  PDA_FCOMB_VARIANT=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:fcomb_tc -s 3 -c 1 -f \
    -o gpurun_out/prof_fcomb_v$v python tools/fcomb_sweep.py 16 > gpurun_out/ncu_fcomb_v$v.log 2>&1
  tail -1 gpurun_out/ncu_fcomb_v$v.log | cut -c1-200
done
