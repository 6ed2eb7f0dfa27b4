echo "fused24: $(ALAC_B200_FUSED=1 python scripts/step_once.py 3600 24 96000 | tail -1)" | tee gpurun_out/fused24.log
TAG=c3b ARGS="3600 24 96000" SRC_KERNELS="enc_search enc_final enc_assemble dec_entropy dec_finish" bash scripts/profile_r02.sh
TAG=c2b ARGS="3600 16 44100" SRC_KERNELS="enc_search enc_final dec_fused" bash scripts/profile_r02.sh
