#!/bin/bash
# Kernel-variant experiment: run a bench command once per library under evacuation_b200/variants/.  bash tools/gpu_variants.sh <tag> <cmd...>
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "base: $(timeout 300 "$@" 2>&1 | tail -1)" | tee -a $OUT/variants.txt
for lib in evacuation_b200/variants/lib_*.so; do
  echo "$(basename $lib): $(EVAC_B200_LIB=$PWD/$lib timeout 300 "$@" 2>&1 | tail -1)" | tee -a $OUT/variants.txt
done
