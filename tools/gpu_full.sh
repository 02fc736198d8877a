#!/bin/bash
# full GPU visit: every -m gpu test, smoke, bench (default build: CTA-pair convs, fp16 no-grad path)
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log > gpurun_out/bench.json; head -c 400 gpurun_out/bench.json; echo
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.json'))
print('infer ms', b['ms_per_step'], 'conv TF', b['roofline']['achieved'], 'frac', b['roofline']['frac'], 'fcomb ms', b['roofline_fcomb']['kernel_ms_per_step'])
t=b['train']; print('train ms', t['ms_per_step'], {k:(round(v['ms_per_step'],3), v['tflops']) for k,v in t['kernels'].items()})
print('src', t['source_train']['ms_per_step'], 'joint', {k:v['ms_per_step'] for k,v in t['joint_fixmatch'].items()})
print('eager', b['eager_gpu_baseline'])
PY
