#!/bin/bash
# N-GPU contract run under torchrun (usage: bash tools/gpu_multi.sh <tag> <N>; gpurun --gpus N) (C2 weak scaling + C5 strong scaling in the same line) + the reference arm launched the same way
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=${2:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step","us_per_step_by_rank")})
print(d["e2e"]); print(d["c5"]); print(d["roofline_pairwise"]["cases"])
PY
tail -3 $OUT/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "ref rc=$?"; cut -c1-400 $OUT/bench_ref_n$N.json
