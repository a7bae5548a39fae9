#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_shard_gpu.py -m gpu -x -q > $OUT/exp3_pytest.log 2>&1
: > $OUT/exp_scan3.jsonl
for TUNE in "scan_sample=0" "scan_sample=64" "scan_sample=16" "scan_sample=256"; do
  for B in 32 128 256 1024; do
    echo "{\"cfg\": \"$TUNE B=$B\"}" >> $OUT/exp_scan3.jsonl
    python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --tune $TUNE --no-encoder --no-cpu-baseline >> $OUT/exp_scan3.jsonl 2>> $OUT/exp_scan3.err
  done
done
python bench.py --steps 5 --warmup 3 --no-encoder --no-cpu-baseline >> $OUT/exp_scan3.jsonl 2>> $OUT/exp_scan3.err
tail -n 3 $OUT/exp3_pytest.log
