#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md, run under gpurun (one GPU).
# 1) plain run must exit 0, 2) launch list (shares), 3) --set full of the top kernels.
set -e
CMD="python bench.py --steps 2 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"enc_|dec_|scan_" -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:enc_search -s 2 -c 1 -o gpurun_out/prof_enc_search $CMD > gpurun_out/ncu_enc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_lane -s 2 -c 1 -o gpurun_out/prof_dec_lane $CMD > gpurun_out/ncu_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dec_output|enc_assemble" -s 4 -c 2 -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-300
# keep gpurun_out under the 64 MiB merge limit: export the pages we read, drop the reports
for r in prof_enc_search prof_dec_lane prof_small; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page source --csv > gpurun_out/$r.source.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
