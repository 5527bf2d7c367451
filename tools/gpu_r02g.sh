#!/bin/bash
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
# pairwise probe occupancy variants
for wpc in 1 2 4; do
  echo "probe wpc=$wpc" >> $OUT/probe.jsonl
  EVAC_PROBE_WPC=$wpc timeout 300 python tools/probe.py >> $OUT/probe.jsonl 2>> $OUT/probe.err
done
cat $OUT/probe.jsonl | cut -c1-1200
# C4 shapes
for shape in 1024x4 512x8; do
  for warm in 0 64 300; do
    echo "shape=$shape warm=$warm" >> $OUT/c4.log
    EVAC_SHAPE_4096=$shape timeout 300 python tools/prof_case.py 256 4096 20 rollout auto $warm >> $OUT/c4.log 2>&1
    EVAC_SHAPE_4096=$shape timeout 300 python tools/prof_case.py 256 4096 20 step auto $warm >> $OUT/c4.log 2>&1
  done
done
cat $OUT/c4.log
EVAC_SHAPE_4096=512x8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cell_list or kernel_shapes or full_size" > $OUT/pytest_512x8.log 2>&1; echo "pytest 512x8 rc=$?"; tail -3 $OUT/pytest_512x8.log
timeout 600 python -m pytest tests -m gpu -q -x -k "checkpoint or capture_does_not or policy_rollout" > $OUT/pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -8 $OUT/pytest_new.log
