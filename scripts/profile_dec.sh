#!/bin/bash
# ncu --set full of the decode kernels (one GPU); exports raw + source pages as CSV (64 MiB copy-back limit)
CMD="python bench.py --steps 1 --warmup 3"
for k in ${KERNELS:-dec_finish dec_entropy}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv > gpurun_out/prof_$k.source.csv 2>/dev/null
  rm -f gpurun_out/prof_$k.ncu-rep
done
