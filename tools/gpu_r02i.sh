#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -x 2>&1 | tail -6
timeout 600 python bench.py --mode train --no-extras --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_train.json
python - <<'PY'
import json
t=json.load(open('gpurun_out/bench_train.json'))
print('train ms', t['ms_per_step'], {k:(round(v['ms_per_step'],3), v['tflops']) for k,v in t['kernels'].items()})
PY
