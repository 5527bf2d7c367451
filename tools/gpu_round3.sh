#!/bin/bash
# One gpurun call at HEAD (round 2, third session): GPU tests, smoke, the contract bench at the driver's flags and at the defaults,
# the reference arm, ncu launch lists (bench, smoke, config-5 loop), ncu full captures of the step kernel and of the tensor-core
# heads kernel, DRAM traffic in the timed regime, sanitizers incl. the policy kernels.
# Usage (repo root on the GPU box): bash tools/gpu_round3.sh [tag]
TAG=${1:-r02w}
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
    print(f, "us/step", round(1e3 * d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3),
          "e2e us", round(1e3 * d["e2e"]["ms_per_step"], 1), "c5", "%.4g" % d["c5"]["value"], "pairwise", round(d["roofline_pairwise"]["frac"], 3))
PY
kill $SMI
timeout 300 python tools/policy_tc_ab.py 8192 65536 1000 > $OUT/policy_tc_ab.jsonl 2>&1; cut -c1-330 $OUT/policy_tc_ab.jsonl
timeout 300 python tools/policy_fused_bench.py 8192 > $OUT/policy_bench.jsonl 2>&1; timeout 300 python tools/policy_fused_bench.py 65536 >> $OUT/policy_bench.jsonl 2>&1; cat $OUT/policy_bench.jsonl
for impl in fused; do timeout 300 python tools/c5_rollout.py 8192 64 graph $impl >> $OUT/c5.jsonl 2>&1; done
timeout 300 python tools/c5_rollout.py 65536 32 graph fused >> $OUT/c5.jsonl 2>&1; cat $OUT/c5.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --no-c5 > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_smoke.csv python __graft_entry__.py smoke > $OUT/smoke_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $OUT/launches_c5.csv python tools/c5_rollout.py 8192 16 eager fused > $OUT/c5_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evac_warp -s 30 -c 2 -o $OUT/prof_step python tools/prof_timed.py 24 1 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/prof_step.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw.py > $OUT/warp_kernel_ncu_raw.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads_tc -s 4 -c 1 -o $OUT/prof_heads_tc python tools/policy_fused_bench.py 8192 > $OUT/ncu_heads.log 2>&1
ncu -i $OUT/prof_heads_tc.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw.py 'pipe_tensor|pipe_tc|pipe_tmem|mem_tensor' > $OUT/heads_tc_ncu_raw.txt 2>&1
ncu -i $OUT/prof_heads_tc.ncu-rep --page details 2>/dev/null > $OUT/heads_tc_ncu_details.txt
timeout 600 ncu --replay-mode application --cache-control none --clock-control none -k regex:evac_warp -s 96 -c 4 \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --csv --page raw --log-file $OUT/timed_app.csv python tools/prof_timed.py 24 1 > $OUT/timed_app.log 2>&1; echo "ncu app rc=$?"
bash tools/gpu_sanitize.sh $TAG/san
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/san_policy.py 300 > $OUT/san/${tool}_policy.txt 2>&1; echo "$tool policy rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/san/${tool}_policy.txt | tail -1)"; done
ls $OUT
