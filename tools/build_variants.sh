#!/bin/bash
# Build kernel-variant libraries into build/variants/ (git-ignored, travels with gpurun): name=extra nvcc flags
# Usage: bash tools/build_variants.sh name1="-DFOO" name2="-DBAR=2" ...
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  EVAC_B200_LIB=$PWD/build/variants/lib_$name.so EVAC_B200_OBJ_TAG=_$name EVAC_B200_NVCC_EXTRA="$flags" python -m evacuation_b200.build --force 2>&1 | tail -1 &
done
wait
for spec in "$@"; do rm -f evacuation_b200/csrc/*_"${spec%%=*}".o; done
ls -la build/variants/
