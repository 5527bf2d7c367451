#!/bin/bash
OUT=gpurun_out/r02j
mkdir -p $OUT
timeout 300 python tools/zerocopy_probe.py >> $OUT/zerocopy.jsonl 2>> $OUT/zerocopy.err
EVAC_B200_LIB=$PWD/build/variants/lib_staged.so timeout 300 python tools/zerocopy_probe.py >> $OUT/zerocopy.jsonl 2>> $OUT/zerocopy.err
cat $OUT/zerocopy.jsonl; tail -3 $OUT/zerocopy.err
