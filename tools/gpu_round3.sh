#!/bin/bash
# One gpurun call at HEAD (round 2, third session): GPU tests, smoke, the contract bench at the driver's flags and at the defaults,
# the reference arm, ncu launch list of the bench and of the config-5 loop, ncu full capture of the tensor-core heads kernel.
# Usage (repo root on the GPU box): bash tools/gpu_round3.sh [tag]
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_k20.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref_k20.json
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; echo "bench k20 rc=$?"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench_k20", "bench"):
    d = json.load(open("$OUT/%s.json" % f))
    print(f, "us/step", round(1e3 * d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3),
          "e2e us", round(1e3 * d["e2e"]["ms_per_step"], 1), "c5", "%.4g" % d["c5"]["value"], d["c5"].get("ms_per_iteration"), "pairwise", round(d["roofline_pairwise"]["frac"], 3))
PY
timeout 300 python tools/policy_fused_bench.py 8192 > $OUT/policy_bench.jsonl 2>&1; timeout 300 python tools/policy_fused_bench.py 65536 >> $OUT/policy_bench.jsonl 2>&1; cat $OUT/policy_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_smoke.csv python __graft_entry__.py smoke > $OUT/smoke_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $OUT/launches_c5.csv python tools/c5_rollout.py 8192 16 nograph fused > $OUT/c5_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads_tc -s 4 -c 1 -o $OUT/prof_heads_tc python tools/policy_fused_bench.py 8192 > $OUT/ncu_heads.log 2>&1
ncu -i $OUT/prof_heads_tc.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw.py 'tensor|tmem|utc|pipe_tc|pipe_uniform' > $OUT/heads_tc_ncu_raw.txt 2>&1
ncu -i $OUT/prof_heads_tc.ncu-rep --page details 2>/dev/null > $OUT/heads_tc_ncu_details.txt
ls $OUT
