#!/bin/bash
# One gpurun call: GPU parity tests, bench, probes, ncu launch list + full captures of the step kernel and the policy kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
timeout 300 python tools/probe.py > $OUT/probe.json 2>&1; cat $OUT/probe.json
timeout 300 python tools/quick_bench.py > $OUT/quick_bench.jsonl 2>&1; cat $OUT/quick_bench.jsonl
timeout 300 python tools/policy_fused_bench.py 8192 > $OUT/policy_bench.jsonl 2>&1; cat $OUT/policy_bench.jsonl
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evac_ -s 8 -c 2 -o $OUT/prof_step python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evac_policy -s 6 -c 2 -o $OUT/prof_policy python tools/policy_fused_bench.py 8192 > $OUT/ncu_policy.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c5.csv python tools/c5_rollout.py 8192 4 graph fused > $OUT/c5_under_ncu.log 2>&1
ls -la $OUT
