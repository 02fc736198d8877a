#!/bin/bash
# ncu --set full of representative CTA-pair conv launches of the SECOND inference step (launch order inside a step:
# 0-1 64->64 resident, 2-4 level 1, 5-10 levels 2-3 (prior net); 11-21 the same for the U-Net; 22-24 up 2, 25-27 up 1,
# 28 192->64, 29-30 64->64 resident)
mkdir -p gpurun_out
BI="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --mode infer"
$BI > gpurun_out/plain_infer.log 2>&1 || { echo "plain infer failed"; tail -5 gpurun_out/plain_infer.log; exit 1; }
for idx in 1 4 7 10 22 25 28; do
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s $((31 + idx)) -c 1 -f -o gpurun_out/prof_conv_pair_$idx $BI > gpurun_out/ncu_full_conv_$idx.log 2>&1
  tail -1 gpurun_out/ncu_full_conv_$idx.log | cut -c1-120
  python tools/ncu_summary.py full gpurun_out/prof_conv_pair_$idx.ncu-rep > gpurun_out/ncu_full_conv_pair_$idx.md 2>&1
  python tools/ncu_summary.py stalls gpurun_out/prof_conv_pair_$idx.ncu-rep 8 >> gpurun_out/ncu_full_conv_pair_$idx.md 2>&1
done
rm -f gpurun_out/prof_conv_pair_{1,4,7,22,25}.ncu-rep
du -sh gpurun_out
