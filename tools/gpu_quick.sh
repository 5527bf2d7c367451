#!/bin/bash
# Quick GPU iteration: parity tests, quick bench, one ncu capture of the C2 step kernel.  bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $OUT/pytest_gpu.log 2>&1; tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/quick_bench.py > $OUT/quick_bench.jsonl 2>&1; cat $OUT/quick_bench.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evac_ -s 8 -c 1 -o $OUT/prof_step python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > $OUT/ncu_full.log 2>&1
