#!/bin/bash
# compute-sanitizer over every kernel family at HEAD: one-warp kernel (EnvBlock), cell list (one CTA), clusters of 4 and 8 CTAs.
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # tool label args...
  tool=$1; label=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 "$@" > $OUT/${tool}_${label}.txt 2>&1
  echo "$tool $label rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}_${label}.txt | tail -1)"
}
run memcheck warp60 python tools/san_case.py 5 60 3
run racecheck warp60 python tools/san_case.py 5 60 3
run memcheck warp33_grav python tools/san_case.py 3 33 3 grav
run memcheck cells4096 python tools/san_case.py 2 4096 2
run racecheck cells1000 python tools/san_case.py 2 1000 2
run memcheck cluster4_9000 python tools/san_case.py 1 9000 2
run racecheck cluster4_9000 python tools/san_case.py 1 9000 1
run synccheck cluster4_9000 python tools/san_case.py 1 9000 1
run memcheck cluster8_20000 python tools/san_case.py 1 20000 1
