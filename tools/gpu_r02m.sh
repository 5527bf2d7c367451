#!/bin/bash
OUT=gpurun_out/r02m
mkdir -p $OUT
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; python -c "
import json; d=json.load(open('$OUT/bench_k20.json')); print('K20', d['ms_per_step'], d['config']['ms_per_step_incl_graph_launch'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['l2_resident']['ms_per_step'], d['rollout']['ms_per_step'], d['c5']['value'], d['c5']['ms_per_iteration'])
for k,v in d['other_workloads'].items():
    print(k, {kk:(round(vv['us_per_step'],2) if isinstance(vv,dict) and 'us_per_step' in vv else None) for kk,vv in v.items() if kk in ('per_step','rollout')}, v.get('ms_per_step'))
"
tail -3 $OUT/bench_k20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_k20.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref_k20.json
