#!/bin/bash
# one full ncu capture of the fused Fcomb kernel inside the inference bench + summaries
mkdir -p gpurun_out
BI="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --mode infer"
$BI > gpurun_out/plain_infer.log 2>&1 || { echo "plain infer failed"; tail -5 gpurun_out/plain_infer.log; exit 1; }
python - <<'PY'
import json
d = json.loads(open('gpurun_out/plain_infer.log').read().strip().splitlines()[-1])
print('infer ms', d['ms_per_step'], 'fcomb ms', d['roofline_fcomb']['kernel_ms_per_step'])
PY
ncu --set full --clock-control none --import-source on -k regex:fcomb_tc -s 2 -c 1 -f -o gpurun_out/prof_fcomb_tc $BI > gpurun_out/ncu_full_fcomb.log 2>&1
tail -1 gpurun_out/ncu_full_fcomb.log | cut -c1-160
python tools/ncu_summary.py full gpurun_out/prof_fcomb_tc.ncu-rep > gpurun_out/ncu_full_fcomb_tc.md 2>&1
python tools/ncu_summary.py stalls gpurun_out/prof_fcomb_tc.ncu-rep 40 >> gpurun_out/ncu_full_fcomb_tc.md 2>&1
rm -f gpurun_out/*.ncu-rep
