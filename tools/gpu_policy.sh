#!/bin/bash
# Policy kernels on the GPU: parity tests, stand-alone timings, config-5 loop, ncu of the embedding kernel.  bash tools/gpu_policy.sh <tag>
TAG=${1:-pol}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_policy.py -x -q > $OUT/pytest_policy.log 2>&1; tail -25 $OUT/pytest_policy.log
timeout 300 python tools/policy_fused_bench.py 8192 > $OUT/policy_bench.jsonl 2>&1; timeout 300 python tools/policy_fused_bench.py 65536 >> $OUT/policy_bench.jsonl 2>&1; cat $OUT/policy_bench.jsonl
for impl in fused torch; do timeout 300 python tools/c5_rollout.py 8192 64 graph $impl >> $OUT/c5.jsonl 2>&1; done
timeout 300 python tools/c5_rollout.py 65536 32 graph fused >> $OUT/c5.jsonl 2>&1; cat $OUT/c5.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evac_policy -s 6 -c 2 -o $OUT/prof_policy python tools/policy_fused_bench.py 8192 > $OUT/ncu_policy.log 2>&1
ls -la $OUT
