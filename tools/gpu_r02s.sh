#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do
  PDA_DEFER_L2_GRADS=$v timeout 600 python bench.py --mode train --no-extras --no-cpu-baseline 2>&1 | grep -v "^frame" | tail -1 > gpurun_out/bench_train_$v.json
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_train_$v.json'))
t = d.get('train', d)
print('defer=$v', 'train ms', round(t['ms_per_step'], 3), 'launches', t.get('gpu_launches'), {k: round(x['ms_per_step'], 3) for k, x in t.get('kernels', {}).items()})
PY
done
timeout 1500 python -m pytest tests/test_gpu_step_differential.py tests/test_gpu_train.py tests/test_gpu_architectures.py tests/test_gpu_ddp.py -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | grep -v "^frame\|^$" | tail -8
