"""Summarise an `ncu --page source --csv` dump: executed warp-instructions by SASS opcode and the top
stall reasons.  Usage: ncu -i X.ncu-rep --page source --csv | python scripts/ncu_source_summary.py"""
import collections
import csv
import sys

lines = sys.stdin.read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
tot = 0
op = collections.Counter()
stall = collections.Counter()
for r in rows:
    try:
        ex = int((r["Instructions Executed"] or "0").replace(",", ""))
    except ValueError:
        continue
    tot += ex
    s = r["Source"].split()
    if s:
        m = s[0] if not s[0].startswith("@") else (s[1] if len(s) > 1 else s[0])
        op[m.split(".")[0]] += ex
    for k, v in r.items():
        if k and k.startswith("stall_") and "Not Issued" not in k and v:
            try:
                stall[k] += int(v.replace(",", ""))
            except ValueError:
                pass
print(f"total executed warp-instructions: {tot}  (SASS lines: {len(rows)})")
for k, v in op.most_common(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    print(f"  {k:12s} {v:12d} {v / tot:6.3f}")
st = sum(stall.values()) or 1
print("stall samples:", ", ".join(f"{k[6:]} {v / st:.2f}" for k, v in stall.most_common(7)))
