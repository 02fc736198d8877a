#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/dbg_up.py 2>&1 | grep -v "^frame" | tail -20
for f in 1 0; do
PDA_FUSE_UPSAMPLE=$f timeout 600 python bench.py --mode infer --no-extras --no-cpu-baseline 2>&1 | grep -v "^frame" | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('fuse=$f infer ms', d['ms_per_step'], 'roof', d['roofline'].get('achieved'), d['roofline'].get('frac'))
"
done
timeout 1500 python -m pytest tests/test_gpu_punet.py tests/test_gpu_baseline_shapes.py tests/test_gpu_conv.py -m gpu -q --timeout 900 -p no:cacheprovider -x 2>&1 | grep -v "^frame\|^$" | tail -15
