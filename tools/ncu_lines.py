"""Per-CUDA-source-line executed warp-instruction totals from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (first function section per file).
Usage: ... | python tools/ncu_lines.py [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
per_line = defaultdict(lambda: [0, 0, ""])
cur_file, cur_line, hdr, seen_funcs = "", None, None, set()
skip = False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        key = (cur_file, r[1])
        skip = key in seen_funcs or (len(seen_funcs) > 0 and r[1] not in {f for _, f in seen_funcs})
        seen_funcs.add(key)
        continue
    if r[0] == "Line No":
        hdr = r
        ix = hdr.index("Instructions Executed")
        isamp = hdr.index("# Samples")
        continue
    if hdr is None or skip or len(r) != len(hdr):
        continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0]))
        per_line[cur_line][2] = r[1].strip()
    elif cur_line is not None and r[2].startswith("0x"):
        per_line[cur_line][0] += int(r[ix] or 0)
        per_line[cur_line][1] += int(r[isamp] or 0)
tot = sum(v[0] for v in per_line.values())
tsamp = sum(v[1] for v in per_line.values())
print("total warp-inst", tot, "samples", tsamp)
for (f, ln), (n, s, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<4d} {n:9d} {100 * n / tot:5.1f}%  samp {100 * s / max(tsamp, 1):5.1f}%  {src[:110]}")
