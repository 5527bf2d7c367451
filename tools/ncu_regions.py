"""Executed warp-instructions of the step kernel grouped by code region (source-line ranges of
evac_kernels.cuh), from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
Usage: ... | python tools/ncu_regions.py <warps-launched> [kernel-substring]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
warps = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
want = sys.argv[2] if len(sys.argv) > 2 else None  # optional kernel-name substring (default: the first kernel of the report)
# region markers: a line `// @region name` in a .cuh starts a region (until the next marker)
OWN = ("evac_kernels.cuh", "evac_warp.cuh")
marks = {f: [(i + 1, m.group(1)) for i, l in enumerate(open("evacuation_b200/csrc/" + f).read().splitlines())
             for m in [re.search(r"@region\s+(\S+)", l)] if m] for f in OWN}


def region_of(fname, line):
    if fname.endswith("sm_100_rt.hpp"):
        return "pairwise"
    if fname.endswith("philox.cuh"):
        return "rng"
    if fname not in marks:
        return "other:" + fname
    name = "helpers"
    for ln, nm in marks[fname]:
        if line >= ln:
            name = nm
    return name


# every SASS address is listed once per level of its inline stack (callee line, call-site line, ...):
# attribute it ONCE, to the region of its call site inside the kernel body (the deepest evac_kernels.cuh line).
addr = {}
KERNEL_FILE = next((r[1].split("/")[-1] for r in rows if r and r[0] == "File Path"), "")  # the kernel body's file comes first
cur_file, cur_line, hdr, first_fn, skip = "", None, None, None, False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if first_fn is None and (want is None or want in r[1]):
            first_fn, KERNEL_FILE = r[1], cur_file
        skip = r[1] != first_fn
        continue
    if r[0] == "Line No":
        hdr = r
        ix, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or skip or len(r) != len(hdr):
        continue
    if r[0] != "":
        cur_line = int(r[0])
    elif cur_line is not None and r[2].startswith("0x"):
        key = ((2 if cur_file == KERNEL_FILE else 1), cur_line) if cur_file in OWN else (0, 0)
        prev = addr.get(r[2])
        if prev is None or key > prev[0]:
            addr[r[2]] = (key, region_of(cur_file, cur_line), int(r[ix] or 0), int(r[isamp] or 0))
inst, samp = defaultdict(int), defaultdict(int)
for _, reg, n, sm in addr.values():
    inst[reg] += n
    samp[reg] += sm
tot, ts = sum(inst.values()), sum(samp.values())
print(f"kernel: {first_fn[:80]}\ntotal warp-inst {tot}  ({tot / warps:.0f} per warp)  samples {ts}")
for k, v in sorted(inst.items(), key=lambda kv: -kv[1]):
    print(f"{k:14s} {v:10d} {100 * v / tot:5.1f}%  {v / warps:7.1f}/warp   stall-samples {100 * samp[k] / max(ts, 1):5.1f}%")
