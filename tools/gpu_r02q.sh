#!/bin/bash
OUT=gpurun_out/r02q
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "paired_walk or cell_list_equals" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for grp in 0 1; do
  for warm in 0 64 300; do
    echo "group=$grp warm=$warm" >> $OUT/c4.log
    EVAC_CELL_GROUP=$grp timeout 300 python tools/prof_case.py 256 4096 20 rollout auto $warm >> $OUT/c4.log 2>&1
  done
  EVAC_CELL_GROUP=$grp timeout 300 python tools/prof_case.py 128 8192 10 rollout auto 64 >> $OUT/c4.log 2>&1
  EVAC_CELL_GROUP=$grp timeout 300 python tools/prof_case.py 1024 1000 20 rollout auto 64 >> $OUT/c4.log 2>&1
  EVAC_CELL_GROUP=$grp timeout 300 python tools/prof_case.py 2048 256 20 rollout auto 64 >> $OUT/c4.log 2>&1
done
cat $OUT/c4.log
