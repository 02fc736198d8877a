#!/bin/bash
# A/B inside ONE box: Fcomb kernel variants built as separate libraries, e.g.
#   python tools/build_variant.py l3off -DFC_L3_MMA=0        (fp32 FMA epilogue of round 1)
#   python tools/build_variant.py ring4 -DFC_RING=4 -DFC_GROUPS=2
# then: gpurun -- 'bash tools/gpu_fcomb_ab.sh l3off ring4'  (the default library is always included; two passes)
mkdir -p gpurun_out
for pass in 1 2; do
for v in default "$@"; do
  if [ $v = default ]; then unset PDA_B200_LIB; else export PDA_B200_LIB=$PWD/probabilistic_domain_adaptation_b200/libpda_b200_$v.so; fi
  timeout 600 python bench.py --mode infer --no-extras --no-cpu-baseline 2>&1 | grep -v "^frame" | tail -1 > gpurun_out/bench_$v.json
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_$v.json'))
print('$v', 'infer ms', round(d['ms_per_step'], 3), 'conv', round(d['roofline']['kernel_ms_per_step'], 3), 'fcomb ms', round(d['roofline_fcomb']['kernel_ms_per_step'], 4))
PY
done
done
