#!/bin/bash
OUT=gpurun_out/r02p
mkdir -p $OUT
for hw in 0 1 2 4; do
  echo "halfwarp unroll=$hw" >> $OUT/probe.jsonl
  EVAC_PROBE_HALFWARP=$hw timeout 300 python - >> $OUT/probe.jsonl 2>> $OUT/probe.err <<PY
import ctypes as C, json, sys
sys.path.insert(0, ".")
from evacuation_b200 import _native as nat
lib = nat.load()
ms, pairs = C.c_float(), C.c_double()
out = {}
for (E, n, reps) in ((4096, 60, 200), (65536, 60, 50), (18944, 64, 200)):
    best = 0.0
    for _ in range(3):
        nat.check(lib.evac_probe_pairwise(0, E, n, reps, C.byref(ms), C.byref(pairs)))
        best = max(best, pairs.value / (ms.value * 1e-3))
    out[f"tflops_E{E}_N{n}"] = round(best * 8 / 1e12, 2)
print(json.dumps(out))
PY
done
cat $OUT/probe.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "logging" 2>&1 | tail -2
