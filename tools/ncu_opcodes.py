"""Summarise an `ncu --page source --csv` dump: executed-instruction histogram by opcode and the
stall reasons (first kernel of the report).  Usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_opcodes.py"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(sys.stdin))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ix, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ix]) for r in data)
tsamp = sum(int(r[isamp]) for r in data)
print(rows[hi - 1][:2] if hi else "", "total warp-inst", tot, "sass lines", len(data), "samples", tsamp)
c, s = Counter(), Counter()
for r in data:
    parts = r[isrc].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    c[op] += int(r[ix])
    s[op] += int(r[isamp])
for op, n in c.most_common(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    print(f"{op:12s} {n:10d} {100 * n / tot:5.1f}%   stall-samples {100 * s[op] / max(tsamp, 1):5.1f}%")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stall reasons (all samples):")
tots = {h: sum(int(r[hdr.index(h)] or 0) for r in data) for h in stall_cols}
for h, v in sorted(tots.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {100 * v / max(tsamp, 1):5.1f}%")
