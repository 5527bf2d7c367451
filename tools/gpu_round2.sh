#!/bin/bash
# One gpurun call at HEAD: GPU tests, smoke, the contract bench at the driver's flags and at the defaults, the reference arm,
# ncu launch list, ncu full capture of the step kernel, DRAM traffic in the timed regime, sanitizers.
# Usage (repo root on the GPU box): bash tools/gpu_round2.sh [tag]
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
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
    print(f, "us/step", round(1e3 * d["ms_per_step"], 3), "incl launch", round(1e3 * d["config"]["ms_per_step_incl_graph_launch"], 3), "frac", round(d["roofline"]["frac"], 3),
          "e2e us", round(1e3 * d["e2e"]["ms_per_step"], 1), "c5", "%.3g" % d["c5"]["value"], "pairwise", round(d["roofline_pairwise"]["frac"], 3),
          "flushed", round(1e3 * d["l2_flushed"]["ms_per_step"], 2), "resident", round(1e3 * d["l2_resident"]["ms_per_step"], 2), "rollout", round(1e3 * d["rollout"]["ms_per_step"], 2))
PY
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --no-c5 > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evac_warp -s 30 -c 2 -o $OUT/prof_step python tools/prof_timed.py 24 1 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/prof_step.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw.py > $OUT/warp_kernel_ncu_raw.txt 2>&1
timeout 600 ncu --replay-mode application --cache-control none --clock-control none -k regex:evac_warp -s 96 -c 4 \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --csv --page raw --log-file $OUT/timed_app.csv python tools/prof_timed.py 24 1 > $OUT/timed_app.log 2>&1; echo "ncu app rc=$?"
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/timed_app.csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
for r in rows[h + 2:]:
    print({k: r[hdr.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active")})
PY
bash tools/gpu_sanitize.sh $TAG/san
ls $OUT
