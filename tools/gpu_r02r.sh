#!/bin/bash
# A/B inside one box: Fcomb variants (fa: fp32 epilogue, late A1 = round 1; fb: fp32 epilogue, early A1;
# fc: last-layer MMA, late A1; default: last-layer MMA, early A1)
mkdir -p gpurun_out
for v in fa fb fc default fa fb fc default; do
  if [ $v = default ]; then unset PDA_B200_LIB; else export PDA_B200_LIB=$PWD/probabilistic_domain_adaptation_b200/libpda_b200_$v.so; fi
  timeout 600 python bench.py --mode infer --no-extras --no-cpu-baseline 2>&1 | grep -v "^frame" | tail -1 > gpurun_out/bench_$v.json
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_$v.json'))
print('$v', 'infer ms', round(d['ms_per_step'], 3), 'conv', round(d['roofline']['kernel_ms_per_step'], 3), 'fcomb ms', round(d['roofline_fcomb']['kernel_ms_per_step'], 4))
PY
done
