#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -x -k "fused_upsample" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_punet.py tests/test_gpu_baseline_shapes.py tests/test_gpu_guards.py -m gpu -q --timeout 600 -p no:cacheprovider -rf 2>&1 | tail -6
for F in 1 0; do
  PDA_FUSE_UPSAMPLE=$F timeout 600 python bench.py --mode infer --no-extras --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_fuse$F.json
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_fuse$F.json'))
print('fuse=$F infer ms', round(b['ms_per_step'],3), 'conv TF', round(b['roofline']['achieved'],1), 'conv ms', round(b['roofline']['kernel_ms_per_step'],3), 'launches', b['gpu_launches'])
PY
done
