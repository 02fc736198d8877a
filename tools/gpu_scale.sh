#!/bin/bash
# N-GPU run on one box: data-parallel gradient check (tools/ddp_check.py) + bench.py --gpus N  (usage: gpu_scale.sh N)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ddp_check.py 2>&1 | tail -3
timeout 1200 $TR bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
b = json.load(open('gpurun_out/bench_n$N.json'))
print('N=$N infer', round(b['value'] / 1e9, 3), 'Gpx*samples/s', round(b['ms_per_step'], 3), 'ms; e2e', round(b['e2e']['value'] / 1e9, 3),
      '; train', b['train_scaling'])
t = b.get('train', {})
print({k: v.get('ms_per_step') for k, v in (t.get('joint_fixmatch') or {}).items()}, (t.get('source_train') or {}).get('ms_per_step'))
PY
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
