#!/bin/bash
# ncu evidence for the shipped build (all ncu runs of one gpurun call): DRAM traffic of the second inference step,
# full captures of the three fused up-sampling conv launches (launch order inside a step: 22 = 768->256 @256^2,
# 25 = 384->128 @512^2, 28 = 192->64 @1024^2) and of the Fcomb kernel.
mkdir -p gpurun_out
BI="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --mode infer"
$BI > gpurun_out/plain_infer.log 2>&1 || { echo "plain infer failed"; tail -5 gpurun_out/plain_infer.log; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv3x3_tc|fcomb_tc" -s 32 -c 32 --csv --log-file gpurun_out/traffic.csv $BI > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-160; wc -l gpurun_out/traffic.csv
for idx in 22 25 28; do
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s $((31 + idx)) -c 1 -f -o gpurun_out/prof_conv_up_$idx $BI > gpurun_out/ncu_full_conv_up_$idx.log 2>&1
  tail -1 gpurun_out/ncu_full_conv_up_$idx.log | cut -c1-120
  python tools/ncu_summary.py full gpurun_out/prof_conv_up_$idx.ncu-rep > gpurun_out/ncu_full_conv_up_$idx.md 2>&1
  python tools/ncu_summary.py stalls gpurun_out/prof_conv_up_$idx.ncu-rep 12 >> gpurun_out/ncu_full_conv_up_$idx.md 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:fcomb_tc -s 2 -c 1 -f -o gpurun_out/prof_fcomb_tc $BI > gpurun_out/ncu_full_fcomb.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_fcomb_tc.ncu-rep > gpurun_out/ncu_full_fcomb_tc.md 2>&1
python tools/ncu_summary.py stalls gpurun_out/prof_fcomb_tc.ncu-rep 25 >> gpurun_out/ncu_full_fcomb_tc.md 2>&1
rm -f gpurun_out/*.ncu-rep
bash tools/gpu_launches.sh
du -sh gpurun_out
