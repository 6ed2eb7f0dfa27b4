"""Hot regions of a kernel from an exported `ncu --page source --csv` file: consecutive SASS instructions with the same
execution count are one region; prints the regions by executed warp instructions with their stall samples.
Usage: ncu_source_hot.py file.source.csv [top_n] [kernel-substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else None
# the file holds one table per kernel: "Kernel Name" row, header row, instruction rows
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        tables.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for t in tables:
    if want and want not in t["name"]:
        continue
    h = {k: i for i, k in enumerate(t["hdr"])}
    stall_cols = [k for k in t["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    ins = []
    for r in t["rows"]:
        try:
            ex = int(float(r[h["Instructions Executed"]] or 0))
        except ValueError:
            continue
        smp = int(float(r[h["# Samples"]] or 0))
        st = {k: int(float(r[h[k]] or 0)) for k in stall_cols}
        ins.append((r[h["Address"]], r[h["Source"]], ex, smp, st))
    total_ex = sum(i[2] for i in ins)
    total_smp = sum(i[3] for i in ins)
    print(f"== {t['name'][:110]}\n   {len(ins)} SASS instructions, {total_ex / 1e6:.1f} M warp instructions executed, {total_smp} samples")
    regions = []
    for i in ins:
        if regions and regions[-1]["ex"] == i[2]:
            g = regions[-1]
        else:
            g = {"start": i[0], "ex": i[2], "n": 0, "smp": 0, "st": {}, "ops": {}}
            regions.append(g)
        g["n"] += 1
        g["smp"] += i[3]
        g["end"] = i[0]
        op = i[1].split()[1] if i[1].startswith("@") else i[1].split()[0]
        g["ops"][op.split(".")[0]] = g["ops"].get(op.split(".")[0], 0) + 1
        for k, v in i[4].items():
            g["st"][k] = g["st"].get(k, 0) + v
    regions.sort(key=lambda g: -g["ex"] * g["n"])
    for g in regions[:top]:
        st = sorted(g["st"].items(), key=lambda kv: -kv[1])[:4]
        ops = sorted(g["ops"].items(), key=lambda kv: -kv[1])[:6]
        print(f"   {g['start'][-5:]}..{g['end'][-5:]} n={g['n']:4d} x {g['ex'] / 1e3:9.0f} k = {g['ex'] * g['n'] / total_ex * 100:5.1f} % of instr, "
              f"{g['smp'] / max(total_smp, 1) * 100:5.1f} % of samples; " + ", ".join(f"{k[6:]} {v}" for k, v in st if v) +
              " | " + " ".join(f"{k}{v}" for k, v in ops))
