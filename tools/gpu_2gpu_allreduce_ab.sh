#!/bin/bash
# 2-GPU check: gradient equality vs single GPU (tools/ddp_check.py), bench with / without SMs reserved for NCCL, bf16 wire
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ddp_check.py 2>&1 | tail -6
for RS in 4 0; do
  timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --reserve-sms $RS 2>/dev/null | tail -1 > gpurun_out/bench_2gpu_rs$RS.json
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_2gpu_rs$RS.json'))
print('reserve $RS: infer', round(b['value']/1e9,3), 'Gpx*samples/s', 'train', round(b['train']['value'],1), 'img/s', round(b['train']['ms_per_step'],3), 'ms', b['train_scaling'])
PY
done
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --mode train --grad-comm bf16 2>/dev/null | tail -1 > gpurun_out/bench_2gpu_bf16wire.json
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_2gpu_bf16wire.json'))
print('bf16 wire: train', round(b['value'],1), 'img/s', round(b['ms_per_step'],3), 'ms')
PY
