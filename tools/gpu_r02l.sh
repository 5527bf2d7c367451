#!/bin/bash
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu --no-c5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; python -c "
import json; d=json.load(open('$OUT/bench_k20.json')); print('K20', d['ms_per_step'], d['config']['ms_per_step_incl_graph_launch'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['config']['timed_by'][:60])"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/bench.json')); print('K2000', d['ms_per_step'], d['config']['ms_per_step_incl_graph_launch'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['c5']['value'], d['cpu_baseline']['value'])"
tail -3 $OUT/bench.err
