"""List the SASS of one region (see tools/ncu_regions.py) with executed counts per warp.
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_region_sass.py <region> <warps> [kernel-substring]"""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
region, warps = sys.argv[1], float(sys.argv[2])
want = sys.argv[3] if len(sys.argv) > 3 else None
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


KERNEL_FILE = next((r[1].split("/")[-1] for r in rows if r and r[0] == "File Path"), "")
addr, order = {}, []
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
        if prev is None:
            order.append(r[2])
        if prev is None or key > prev[0]:
            addr[r[2]] = (key, region_of(cur_file, cur_line), int(r[ix] or 0), int(r[isamp] or 0), r[3], cur_file, cur_line)
for ad in sorted(addr):
    key, reg, n, sm, sass, f, ln = addr[ad]
    if reg == region and n:
        print(f"{ad[-5:]} {n / warps:6.2f} s{sm:3d} {f[:12]}:{ln:<4d} {sass.strip()}")
