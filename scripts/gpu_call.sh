python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo n8 rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["encode_msamples_s"], d["decode_msamples_s"], d["x_all_host_cores"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["e2e_over_floor"])
print(d["call_ms_per_rank"])
c=d["config3"]
print("c3", c["value"], c["ms_per_step"], c["encode_msamples_s"], c["decode_msamples_s"], c["x_all_host_cores"], c["cpu_baseline"]["value"], c["kernel_ms_per_step"])
print(c["call_ms_per_rank"])
PY
