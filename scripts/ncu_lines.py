"""Executed warp-instructions and stall samples per CUDA source line from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--kernel-id ...]`.
Usage: ... | python scripts/ncu_lines.py [top_n]"""
import collections
import csv
import sys

lines = sys.stdin.read().splitlines()
inst = collections.Counter()
samp = collections.Counter()
text = {}
i = 0
while i < len(lines):
    if lines[i].startswith('"Line No"'):
        hdr = next(csv.reader([lines[i]]))
        li, ie, sa = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        i += 1
        while i < len(lines) and not lines[i].startswith('"File Path"'):
            r = next(csv.reader([lines[i]]))
            i += 1
            if len(r) <= ie or not r[li].strip().isdigit():
                continue
            ln = int(r[li])
            text.setdefault(ln, r[1].strip())
            try:
                inst[ln] += int((r[ie] or "0").replace(",", ""))
                samp[ln] += int((r[sa] or "0").replace(",", ""))
            except ValueError:
                pass
    else:
        i += 1
tot, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp-instructions {tot}, stall samples {ts}")
for ln, v in inst.most_common(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    print(f"{ln:5d} inst {v / tot:6.3f} samp {samp[ln] / ts:6.3f}  {text[ln][:100]}")
