#!/bin/bash
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
for rep in 1 2; do
for v in old "" obsdirect nopair floor; do
  if [ -n "$v" ]; then export EVAC_B200_LIB=$PWD/build/variants/lib_$v.so; else unset EVAC_B200_LIB; fi
  echo "variant=$v" >> $OUT/step_bench.jsonl
  timeout 300 python tools/step_bench.py 4096 960 24 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
done
done
unset EVAC_B200_LIB
cat $OUT/step_bench.jsonl
