for w in "5,4,3,2,1" "4,4,3,2,1" "3,3,3,2,1" "6,5,4,2,1" "4,3,2,1" "8,6,4,3,2,1" "3,3,2,1" "2,2,2,1" "3,2,1" "10,8,6,4,2,1,1"; do
  echo "weights=$w $(ALAC_B200_PIPELINE_WEIGHTS=$w python scripts/e2e_probe.py 2>&1 | grep enc_wall | tail -2 | cut -c1-42 | tr '\n' ' ')"
done
