#!/bin/bash
TAG=${1:-exp2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for u in 4 8 2; do echo "unroll $u"; EVAC_PROBE_UNROLL=$u timeout 300 python tools/probe.py > $OUT/probe_u$u.json 2>&1; cat $OUT/probe_u$u.json; done
