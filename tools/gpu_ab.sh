#!/bin/bash
# Same-box A/B of kernel variants in the headline regime (tools/step_bench.py): build the variants first with
#   bash tools/build_variants.sh name1="-DFLAG" name2="-DOTHER=2"      (-> build/variants/lib_<name>.so, travels with gpurun)
# then on the GPU box:   bash tools/gpu_ab.sh <tag> "" name1 name2 ...   ("" = the default library)
# Round-2 variants: -DEVAC_PROBE_EMPTY / _FLOOR / _NOPAIR / _NOOBS / _NOSTATE (phases compiled out), -DEVAC_OBS_STAGED,
# -DEVAC_WARP_MINB=32, -DEVAC_PAIR_UNROLL=2|8; runtime switches: EVAC_WARP_WPC, EVAC_SHAPE_4096, EVAC_CELL_DYNAMIC, EVAC_CELL_GROUP.
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "$@"; do
  if [ -n "$v" ]; then export EVAC_B200_LIB=$PWD/build/variants/lib_$v.so; else unset EVAC_B200_LIB; fi
  echo "variant=$v" >> $OUT/step_bench.jsonl
  timeout 300 python tools/step_bench.py 4096 960 24 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
done
unset EVAC_B200_LIB
cat $OUT/step_bench.jsonl
