"""Loops of a kernel from `cuobjdump -sass`: for every backward branch, the instruction count and opcode histogram
of the body [target, branch].  Usage: sass_loops.py file.sass [min_instructions]"""
import collections
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\*", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
min_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
print(f"{len(ins)} instructions")
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
for lo, hi in sorted(loops):
    n = hi - lo + 1
    if n < min_n:
        continue
    hist = collections.Counter()
    for _, t in ins[lo:hi + 1]:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        hist[op.split(".")[0]] += 1
    inner = [(l, h) for l, h in loops if l >= lo and h <= hi and (l, h) != (lo, hi) and h - l + 1 >= min_n]
    print(f"loop {ins[lo][0]:#06x}..{ins[hi][0]:#06x}: {n} instr" + (f" (contains {len(inner)} inner loops)" if inner else ""))
    print("   " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
