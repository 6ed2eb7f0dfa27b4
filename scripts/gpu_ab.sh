# developer A/B on the GPU box: step timings of every built library variant (alac_b200/csrc/libalac_b200*.so)
for lib in alac_b200/csrc/libalac_b200*.so; do
  for cfg in "3600 16 44100" "3600 24 96000" ${FULL:+"36000 24 96000"}; do
    echo "$(basename $lib) $cfg: $(ALAC_B200_LIB=$PWD/$lib python scripts/step_once.py $cfg 2>&1 | tail -1)"
  done
done | tee gpurun_out/ab.log
