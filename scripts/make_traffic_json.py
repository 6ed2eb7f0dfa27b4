"""profiles/r01_traffic.json from the exported `ncu --page raw --csv` files of scripts/profile_r01.sh
(per launch: DRAM bytes, duration, issue-slot utilisation).  bench.py reads it for roofline.traffic."""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3}
out = {"source": "ncu --set full --clock-control none, scripts/profile_r01.sh, bench.py workload (1-hour 16-bit/44.1 kHz stereo)",
       "kernels": []}
for path in sorted(glob.glob(os.path.join(ROOT, sys.argv[1] if len(sys.argv) > 1 else "gpurun_out", "prof_*.raw.csv"))):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
        num = lambda k: float(d[k][1].replace(",", "")) * UNIT.get(d[k][0], 1)
        out["kernels"].append({"kernel": d["Kernel Name"][1].split("(")[0], "dram_read_bytes": num("dram__bytes_read.sum"),
                               "dram_write_bytes": num("dram__bytes_write.sum"), "duration_ms": num("gpu__time_duration.sum"),
                               "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                               "warp_instructions": num("smsp__inst_executed.sum")})
json.dump(out, open(os.path.join(ROOT, "profiles", "r01_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
