#!/bin/bash
# Round-2 ncu recipe (/opt/skills/guides/B200_PROFILING.md), one GPU, under gpurun.
#   TAG=c2    ARGS="3600 16 44100"    the bench workload (1 h 16/44.1)
#   TAG=c3_1h ARGS="3600 24 96000"    one hour of the north-star corpus (24/96): same kernels, 1/10 of the packets
#   TAG=c3 LIGHT=1 ARGS="36000 24 96000"   the whole 10-hour corpus: launch list + DRAM bytes / instructions only
# 1) plain run must exit 0, 2) launch list of one step, 3) --set full of every hot kernel of that step (LIGHT=1: four
# metrics instead; --set full replays a 20 GB working set ~40 times).  Reports are exported to CSV here (64 MiB
# copy-back limit) and dropped.
set -e
TAG=${TAG:-c2}
ARGS=${ARGS:-3600 16 44100}
OUT=gpurun_out/r02_$TAG
mkdir -p $OUT
CMD="python scripts/step_once.py $ARGS"
K='regex:enc_search|enc_final|enc_assemble|dec_fused|dec_entropy|dec_finish'
$CMD > $OUT/plain.json 2> $OUT/plain.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
if [ -n "$LIGHT" ]; then
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size \
      --clock-control none --profile-from-start off -k "$K" -o $OUT/step $CMD > $OUT/ncu_light.log 2>&1
else
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -o $OUT/step $CMD > $OUT/ncu_full.log 2>&1
fi
ncu -i $OUT/step.ncu-rep --page raw --csv > $OUT/step.raw.csv 2>/dev/null
if [ -z "$LIGHT" ]; then
  for k in ${SRC_KERNELS:-enc_search enc_final enc_assemble dec_fused dec_entropy dec_finish}; do
    ncu -i $OUT/step.ncu-rep --page source --csv -k regex:$k > $OUT/$k.source.csv 2>/dev/null || true
  done
fi
rm -f $OUT/step.ncu-rep
cat $OUT/plain.json
