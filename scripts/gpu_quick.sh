# developer loop on the GPU box: parity tests, then the step timings of the two bench corpora
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_configs.py tests/test_gpu_fork_pin.py -x -q > gpurun_out/quick_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/quick_pytest.log
python scripts/step_once.py 3600 16 44100 | tee gpurun_out/quick_c2.json
python scripts/step_once.py 3600 24 96000 | tee gpurun_out/quick_c3_1h.json
python scripts/step_once.py 36000 24 96000 | tee gpurun_out/quick_c3.json
