"""Static SASS instruction count per kernel of the built library (no GPU needed).
Usage: python tools/sass_count.py [substring-filter]"""
import re
import subprocess
import sys
from collections import Counter

sys.path.insert(0, ".")
from evacuation_b200.build import LIB_PATH

out = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True).stdout
flt = sys.argv[1] if len(sys.argv) > 1 else ""
cur, counts, ops = None, Counter(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        ops[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur] += 1
        ops[cur][m.group(1)] += 1
for k, v in counts.items():
    if flt in k:
        print(v, k)
        print("   ", ", ".join(f"{o}:{n}" for o, n in ops[k].most_common(18)))
