#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md, run under gpurun (one GPU).
# 1) plain run must exit 0, 2) launch list (shares), 3) --set full of every kernel of the step.
set -e
CMD="python bench.py --steps 2 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"enc_|dec_|scan_" -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
for k in enc_search_split enc_final enc_assemble dec_fused; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  # keep gpurun_out under the 64 MiB merge limit: export the pages we read, drop the report
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv > gpurun_out/prof_$k.source.csv 2>/dev/null
  rm -f gpurun_out/prof_$k.ncu-rep
done
tail -1 gpurun_out/plain.log | cut -c1-300
