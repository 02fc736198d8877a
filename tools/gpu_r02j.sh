#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_guards.py tests/test_gpu_train.py tests/test_gpu_conv.py tests/test_gpu_steps.py tests/test_gpu_step_differential.py -m gpu -q --timeout 300 -p no:cacheprovider -rf 2>&1 | tail -12
timeout 600 python bench.py --mode train --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_train.json
python - <<'PY'
import json
t=json.load(open('gpurun_out/bench_train.json'))
print('train ms', t['ms_per_step'], 'python-launched', t['ms_per_step_python_launched'], 'launches', t['gpu_launches'], {k:(round(v['ms_per_step'],3), v['tflops']) for k,v in t['kernels'].items()})
print('src', t['source_train']['ms_per_step'], 'joint', {k:(v['ms_per_step'], v['launches_per_step']) for k,v in t['joint_fixmatch'].items()})
PY
