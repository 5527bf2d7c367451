#!/bin/bash
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "teacher_forced or kernel_shapes or full_size" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
for v in "" floor floor_noobs floor_nostate floor_loadonly; do
  if [ -n "$v" ]; then export EVAC_B200_LIB=$PWD/build/variants/lib_$v.so; else unset EVAC_B200_LIB; fi
  echo "variant=$v" >> $OUT/step_bench.jsonl
  timeout 300 python tools/step_bench.py 4096 960 24 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
done
unset EVAC_B200_LIB
echo "E sweep" >> $OUT/step_bench.jsonl
timeout 300 python tools/step_bench.py 2048 960 48 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
timeout 300 python tools/step_bench.py 8192 480 12 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
timeout 300 python tools/step_bench.py 16384 240 6 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
timeout 300 python tools/step_bench.py 65536 96 2 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
cat $OUT/step_bench.jsonl
