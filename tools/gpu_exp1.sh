#!/bin/bash
# A/B: rotating-set bench line with and without programmatic dependent launch; probe; ncu of the pairwise probe.
TAG=${1:-exp1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for pdl in 0 1; do
  EVAC_PDL=$pdl timeout 300 python bench.py --no-cpu --no-extra > $OUT/bench_pdl$pdl.json 2> $OUT/bench_pdl$pdl.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_pdl$pdl.json"))
print("pdl=$pdl value %.3e us/step %.2f flushed %.2f resident %.2f rollout %.2f e2e %.1f frac %.3f" % (d["value"], d["ms_per_step"]*1e3, d["l2_flushed"]["ms_per_step"]*1e3, d["l2_resident"]["ms_per_step"]*1e3, d["rollout"]["ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["frac"]))
PY
done
tail -3 $OUT/bench_pdl1.err
EVAC_PDL=1 timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_pdl1.log 2>&1; tail -2 $OUT/pytest_pdl1.log
timeout 300 python tools/probe.py > $OUT/probe.json 2>&1; cat $OUT/probe.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:probe_pairwise -s 2 -c 1 -o $OUT/prof_probe python tools/probe.py > $OUT/ncu_probe.log 2>&1; tail -2 $OUT/ncu_probe.log
