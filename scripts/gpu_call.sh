python bench.py > gpurun_out/bench_ov.json 2> gpurun_out/bench_ov.err; echo rc=$?; tail -3 gpurun_out/bench_ov.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ov.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d.get("overlapped_steps"), d["roofline"]["issue_frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
