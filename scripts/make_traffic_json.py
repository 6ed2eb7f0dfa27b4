"""profiles/<out>.json from an exported `ncu --page raw --csv` file (per launch: DRAM bytes, duration, issue-slot
utilisation, warp instructions, L1 wavefront utilisation).  bench.py reads profiles/r02_traffic_<config>.json for
roofline.traffic / roofline.issue_frac.
    python scripts/make_traffic_json.py <raw.csv> <out.json> "<what was captured>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3}
src, dst, what = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
out = {"source": what, "kernels": []}
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, r)}

    def num(k):
        if k not in d or d[k][1] in ("", "n/a"):
            return None
        return float(d[k][1].replace(",", "")) * UNIT.get(d[k][0], 1)
    out["kernels"].append({"kernel": d["Kernel Name"][1].split("(")[0].replace("void ", ""), "dram_read_bytes": num("dram__bytes_read.sum"),
                           "dram_write_bytes": num("dram__bytes_write.sum"), "duration_ms": num("gpu__time_duration.sum"),
                           "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           "l1tex_wavefront_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                           "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                           "registers": num("launch__registers_per_thread"), "grid": num("launch__grid_size"),
                           "inst_executed": num("smsp__inst_executed.sum")})
# several launches of one kernel in the step (a 10-hour corpus runs its encode in scratch-sized chunks): one entry per
# kernel, sums over its launches; percentages are duration-weighted means
agg = {}
for k in out["kernels"]:
    a = agg.setdefault(k["kernel"], {"kernel": k["kernel"], "launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "duration_ms": 0.0,
                                     "inst_executed": 0.0, "registers": k["registers"], "grid": 0.0, "_w": {}})
    a["launches"] += 1
    for f in ("dram_read_bytes", "dram_write_bytes", "duration_ms", "inst_executed", "grid"):
        a[f] += k[f] or 0.0
    for f in ("issue_active_pct", "l1tex_wavefront_pct", "warps_active_pct"):
        if k[f] is not None:
            a["_w"][f] = a["_w"].get(f, 0.0) + k[f] * (k["duration_ms"] or 0.0)
for a in agg.values():
    for f, v in a.pop("_w").items():
        a[f] = round(v / a["duration_ms"], 2) if a["duration_ms"] else None
    a["issue_slot_frac"] = round(a["inst_executed"] / (148 * 4 * 1.965e9 * a["duration_ms"] / 1e3), 4) if a["duration_ms"] else None
out["kernels"] = list(agg.values())
json.dump(out, open(os.path.join(ROOT, dst), "w"), indent=1)
print(json.dumps(out, indent=1))
