#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/exp_scan2.jsonl
for KBS in 1 2 3 6; do
  for B in 32 128 256; do
    echo "{\"cfg\": \"KBS=$KBS B=$B\"}" >> $OUT/exp_scan2.jsonl
    ICD_SCAN_KBS=$KBS python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --no-encoder --no-cpu-baseline >> $OUT/exp_scan2.jsonl 2>> $OUT/exp_scan2.err
  done
done
python bench.py --steps 5 --warmup 3 --no-encoder --no-cpu-baseline >> $OUT/exp_scan2.jsonl 2>> $OUT/exp_scan2.err
