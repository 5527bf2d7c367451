#!/bin/bash
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "" floor empty; do
  for wpc in 1 4 8; do
    if [ -n "$v" ]; then export EVAC_B200_LIB=$PWD/build/variants/lib_$v.so; else unset EVAC_B200_LIB; fi
    echo "variant=$v wpc=$wpc" >> $OUT/step_bench.jsonl
    EVAC_WARP_WPC=$wpc timeout 300 python tools/step_bench.py 4096 960 24 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
  done
done
unset EVAC_B200_LIB
cat $OUT/step_bench.jsonl
