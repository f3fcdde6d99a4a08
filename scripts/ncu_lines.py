"""Executed warp-instructions and stall samples per CUDA source line from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--kernel-id ...]`.
Usage: ... | python scripts/ncu_lines.py [top_n] [file:lo-hi=label ...]   (labels aggregate line ranges)"""
import collections
import csv
import os
import sys

lines = sys.stdin.read().splitlines()
inst = collections.Counter()
samp = collections.Counter()
text = {}
cur = "?"
i = 0
while i < len(lines):
    if lines[i].startswith('"File Path"'):
        cur = os.path.basename(next(csv.reader([lines[i]]))[1])
        i += 1
    elif lines[i].startswith('"Line No"'):
        hdr = next(csv.reader([lines[i]]))
        li, ie, sa = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        i += 1
        while i < len(lines) and not lines[i].startswith('"File Path"'):
            r = next(csv.reader([lines[i]]))
            i += 1
            if len(r) <= ie or not r[li].strip().isdigit():
                continue
            key = (cur, int(r[li]))
            text.setdefault(key, r[1].strip())
            try:
                inst[key] += int((r[ie] or "0").replace(",", ""))
                samp[key] += int((r[sa] or "0").replace(",", ""))
            except ValueError:
                pass
    else:
        i += 1
tot, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp-instructions {tot}, stall samples {ts}")
args = sys.argv[1:]
topn = int(args[0]) if args and args[0].isdigit() else 30
for spec in [a for a in args if "=" in a and ":" in a]:
    rng, label = spec.split("=")
    f, lohi = rng.split(":")
    lo, hi = (int(x) for x in lohi.split("-"))
    v = sum(c for (ff, ln), c in inst.items() if ff.startswith(f) and lo <= ln <= hi)
    s = sum(c for (ff, ln), c in samp.items() if ff.startswith(f) and lo <= ln <= hi)
    print(f"  region {label:28s} inst {v / tot:6.3f} samp {s / ts:6.3f}")
byfile = collections.Counter()
for (f, ln), v in inst.items():
    byfile[f] += v
print("by file:", {f: round(v / tot, 3) for f, v in byfile.items()})
for key, v in inst.most_common(topn):
    print(f"{key[0][:18]:18s}{key[1]:5d} inst {v / tot:6.3f} samp {samp[key] / ts:6.3f}  {text[key][:90]}")
