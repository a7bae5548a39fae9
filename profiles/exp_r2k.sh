#!/bin/bash
# r02k: stride of the (now cheap) slot-maxima pre-pass
OUT=gpurun_out; mkdir -p $OUT
for cfg in "12500000 1024 40 30" "12500000 256 150 100" "12500000 128 200 150" "100000000 128 40 30" "100000000 1024 10 8" "1000000 256 400 300"; do
  set -- $cfg
  ROWS=$1 BATCH=$2 STEPS=$3 WARM=$4 VARIANTS="scan_sample=256;scan_sample=64;scan_sample=32;scan_sample=16;scan_sample=8" timeout 900 python profiles/scan_ab.py >> $OUT/r02k_stride_ab.jsonl 2>> $OUT/r02k_stride_ab.err
done
cat $OUT/r02k_stride_ab.jsonl; tail -3 $OUT/r02k_stride_ab.err
