#!/bin/bash
OUT=gpurun_out/r02k
mkdir -p $OUT
for dyn in 0 1; do
  for warm in 0 64 300; do
    echo "dynamic=$dyn warm=$warm" >> $OUT/c4.log
    EVAC_CELL_DYNAMIC=$dyn timeout 300 python tools/prof_case.py 256 4096 20 rollout auto $warm >> $OUT/c4.log 2>&1
  done
  EVAC_CELL_DYNAMIC=$dyn timeout 300 python tools/prof_case.py 128 8192 10 rollout auto 64 >> $OUT/c4.log 2>&1
  EVAC_CELL_DYNAMIC=$dyn timeout 300 python tools/prof_case.py 1024 1000 20 rollout auto 64 >> $OUT/c4.log 2>&1
done
cat $OUT/c4.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cell_list or kernel_shapes or full_size or c4_flocked or cluster_pass" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu --no-c5 > $OUT/bench_k20.json 2> $OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/bench_k20.json')); print(d['ms_per_step'], d['config']['ms_per_step_incl_graph_launch'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
