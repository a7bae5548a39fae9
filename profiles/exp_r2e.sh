#!/bin/bash
# r02e: would a tile-major table layout ([tile][K block][128 rows][64]: 8 KiB contiguous runs instead of 768-byte row
# segments) buy HBM efficiency under the power cap?  Timing only (profiling build reads the row-major bytes AS IF tiled).
OUT=gpurun_out; mkdir -p $OUT
export PROF_LIB=profiles/_prof/libicdrag_prof.so
for cfg in "10000000 128 200 150" "10000000 256 150 100" "12500000 1024 40 30" "10000000 32 200 150"; do
  set -- $cfg
  ROWS=$1 BATCH=$2 STEPS=$3 WARM=$4 VARIANTS="scan_tiled=0;scan_tiled=1" timeout 600 python profiles/scan_ab.py >> $OUT/r02e_tiled_ab.jsonl 2>> $OUT/r02e_tiled_ab.err
done
cat $OUT/r02e_tiled_ab.jsonl; tail -3 $OUT/r02e_tiled_ab.err
