#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench_k20.json; tail -5 $OUT/bench_k20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_k20.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref_k20.json
