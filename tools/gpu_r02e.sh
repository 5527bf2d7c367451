#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
for v in "" staged minb32 unr2 unr8; do
  if [ -n "$v" ]; then export EVAC_B200_LIB=$PWD/build/variants/lib_$v.so; else unset EVAC_B200_LIB; fi
  echo "variant=$v" >> $OUT/step_bench.jsonl
  timeout 300 python tools/step_bench.py 4096 960 24 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
done
unset EVAC_B200_LIB
for wpc in 2 4; do
  echo "wpc=$wpc" >> $OUT/step_bench.jsonl
  EVAC_WARP_WPC=$wpc timeout 300 python tools/step_bench.py 4096 960 24 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
done
echo "grav" >> $OUT/step_bench.jsonl
timeout 300 python tools/step_bench.py 4096 960 24 grav >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
cat $OUT/step_bench.jsonl
