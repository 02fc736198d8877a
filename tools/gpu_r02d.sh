#!/bin/bash
mkdir -p gpurun_out
PDA_DEBUG=1 timeout 600 python tools/conv_layer_bench.py 2>&1 | tee gpurun_out/conv_layer_bench.md
PDA_CONV_PAIR=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --mode infer > gpurun_out/plain_pair.log 2>&1 &&
PDA_CONV_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 3 -c 3 -f -o gpurun_out/prof_conv_pair python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --mode infer > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log
ls -la gpurun_out | tail -5
