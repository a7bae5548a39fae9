#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_scan_gpu.py -m gpu -x -q -k knobs > $OUT/exp4_pytest.log 2>&1
: > $OUT/exp_scan4.jsonl
for S in 128 256 512 1024 2048; do
  for B in 8 128 1024; do
    echo "{\"cfg\": \"10M sample=$S B=$B\"}" >> $OUT/exp_scan4.jsonl
    python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --tune scan_sample=$S --no-encoder --no-cpu-baseline >> $OUT/exp_scan4.jsonl 2>> $OUT/exp_scan4.err
  done
done
for S in 256 1024 4096; do
  for B in 128 1024; do
    echo "{\"cfg\": \"100M sample=$S B=$B\"}" >> $OUT/exp_scan4.jsonl
    python bench.py --batch $B --steps 5 --warmup 3 --tune scan_sample=$S --no-encoder --no-cpu-baseline >> $OUT/exp_scan4.jsonl 2>> $OUT/exp_scan4.err
  done
done
tail -n 3 $OUT/exp4_pytest.log
