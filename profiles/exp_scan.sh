#!/bin/bash
# scan-kernel experiments: BN / drift limiter / query tiles per launch, 10 M rows
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/exp_scan.jsonl
for cfg in "128 0 32" "128 4 8" "64 0 32" "64 4 8" "64 8 8"; do
  set -- $cfg
  for B in 128 1024 4096; do
    echo "{\"cfg\": \"BN=$1 DRIFT=$2 TMAX=$3 B=$B\"}" >> $OUT/exp_scan.jsonl
    ICD_SCAN_BN=$1 ICD_SCAN_DRIFT=$2 ICD_SCAN_TMAX=$3 python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --no-encoder --no-cpu-baseline >> $OUT/exp_scan.jsonl 2>> $OUT/exp_scan.err
  done
done
ICD_SCAN_BN=64 timeout 600 python -m pytest tests/test_scan_gpu.py -m gpu -x -q > $OUT/exp_pytest_bn64.log 2>&1
timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_shard_gpu.py -m gpu -x -q > $OUT/exp_pytest_bn128.log 2>&1
tail -3 $OUT/exp_pytest_bn64.log $OUT/exp_pytest_bn128.log
