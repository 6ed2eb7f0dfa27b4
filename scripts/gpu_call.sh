for lib in alac_b200/csrc/libalac_b200.so alac_b200/csrc/libalac_b200_bulk.so; do
  for cfg in "3600 16 44100" "3600 24 96000"; do
    echo "$(basename $lib) $cfg: $(ALAC_B200_LIB=$PWD/$lib timeout 120 python scripts/step_once.py $cfg 2>&1 | tail -1)"
  done
done | tee gpurun_out/ab_bulk.log
ALAC_B200_LIB=$PWD/alac_b200/csrc/libalac_b200_bulk.so timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k regex:enc_final2 --csv python scripts/step_once.py 3600 24 96000 > gpurun_out/ab_bulk_ncu.csv 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k regex:enc_final2 --csv python scripts/step_once.py 3600 24 96000 > gpurun_out/ab_base_ncu.csv 2>&1
grep -h "enc_final2" gpurun_out/ab_bulk_ncu.csv gpurun_out/ab_base_ncu.csv | cut -d, -f5,13-15 | head -20
