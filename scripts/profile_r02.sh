#!/bin/bash
# Round-2 ncu recipe (/opt/skills/guides/B200_PROFILING.md), one GPU, under gpurun.
#   TAG=c2 ARGS="3600 16 44100"   the bench workload (1 h 16/44.1)
#   TAG=c3 ARGS="3600 24 96000"   one hour of the north-star corpus (24/96): same kernels, 1/10 of the packets
# 1) plain run must exit 0, 2) launch list of one step, 3) --set full of every hot kernel of that step; reports are
# exported to CSV here (64 MiB copy-back limit) and dropped.
set -e
TAG=${TAG:-c2}
ARGS=${ARGS:-3600 16 44100}
OUT=gpurun_out/r02_$TAG
mkdir -p $OUT
CMD="python scripts/step_once.py $ARGS"
$CMD > $OUT/plain.json 2> $OUT/plain.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"enc_search|enc_final|enc_assemble|dec_fused|dec_entropy|dec_finish" -o $OUT/step $CMD > $OUT/ncu_full.log 2>&1
ncu -i $OUT/step.ncu-rep --page raw --csv > $OUT/step.raw.csv 2>/dev/null
for k in ${SRC_KERNELS:-enc_search enc_final dec_fused dec_entropy dec_finish}; do
  ncu -i $OUT/step.ncu-rep --page source --csv -k regex:$k > $OUT/$k.source.csv 2>/dev/null || true
done
rm -f $OUT/step.ncu-rep
cat $OUT/plain.json
