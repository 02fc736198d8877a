#!/bin/bash
# ncu evidence for the shipped build (all ncu runs of one gpurun call count as one): DRAM traffic of one inference step,
# full captures of the conv (CTA-pair), fused Fcomb and weight-gradient kernels.  Summaries are made on the box; the
# .ncu-rep files are only kept while gpurun_out stays below the 64 MiB that travel back.
mkdir -p gpurun_out
BI="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --mode infer"
BT="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-graph --mode train"
$BI > gpurun_out/plain_infer.log 2>&1 || { echo "plain infer failed"; tail -5 gpurun_out/plain_infer.log; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv3x3_tc|fcomb_tc" -s 32 -c 32 --csv --log-file gpurun_out/traffic.csv $BI > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-160; wc -l gpurun_out/traffic.csv
i=0
for pat in "conv3x3_tc2_kernel<256" "conv3x3_tc2_kernel<128" "conv3x3_tc2_kernel<64, 2, 1" "conv3x3_tc2_kernel<64, 2, 0"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 4 -c 2 -f -o gpurun_out/prof_conv_pair_$i $BI > gpurun_out/ncu_full_conv_$i.log 2>&1
  tail -1 gpurun_out/ncu_full_conv_$i.log | cut -c1-160
  python tools/ncu_summary.py full gpurun_out/prof_conv_pair_$i.ncu-rep > gpurun_out/ncu_full_conv_pair_$i.md 2>&1
  python tools/ncu_summary.py stalls gpurun_out/prof_conv_pair_$i.ncu-rep 12 >> gpurun_out/ncu_full_conv_pair_$i.md 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:fcomb_tc -s 2 -c 1 -f -o gpurun_out/prof_fcomb_tc $BI > gpurun_out/ncu_full_fcomb.log 2>&1
tail -1 gpurun_out/ncu_full_fcomb.log | cut -c1-160
python tools/ncu_summary.py full gpurun_out/prof_fcomb_tc.ncu-rep > gpurun_out/ncu_full_fcomb_tc.md 2>&1
python tools/ncu_summary.py stalls gpurun_out/prof_fcomb_tc.ncu-rep 15 >> gpurun_out/ncu_full_fcomb_tc.md 2>&1
$BT > gpurun_out/plain_train.log 2>&1 || { echo "plain train failed"; tail -5 gpurun_out/plain_train.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:wgrad3x3 -s 60 -c 4 -f -o gpurun_out/prof_wgrad $BT > gpurun_out/ncu_full_wgrad.log 2>&1
tail -1 gpurun_out/ncu_full_wgrad.log | cut -c1-160
python tools/ncu_summary.py full gpurun_out/prof_wgrad.ncu-rep > gpurun_out/ncu_full_wgrad.md 2>&1
python tools/ncu_summary.py stalls gpurun_out/prof_wgrad.ncu-rep 12 >> gpurun_out/ncu_full_wgrad.md 2>&1
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
# keep the reports only if everything fits
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_conv_pair_[234].ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/*.ncu-rep; fi
du -sh gpurun_out
