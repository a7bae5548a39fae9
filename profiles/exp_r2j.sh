#!/bin/bash
# r02j: slot-maxima pre-pass: parity, then same-box A/B against the list-based pre-pass at the 8-GPU shard size and 10 M rows
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_scan_gpu.py -m gpu -x -q ) > $OUT/r02j_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 $OUT/r02j_pytest.log
for cfg in "12500000 1024 40 30" "12500000 256 150 100" "12500000 128 200 150" "10000000 32 200 150" "10000000 4096 12 8"; do
  set -- $cfg
  ROWS=$1 BATCH=$2 STEPS=$3 WARM=$4 VARIANTS="scan_pre_slots=0;scan_pre_slots=1;scan_pre_slots=1,scan_sample=64;scan_pre_slots=1,scan_sample=1024" timeout 600 python profiles/scan_ab.py >> $OUT/r02j_prepass_ab.jsonl 2>> $OUT/r02j_prepass_ab.err
done
cat $OUT/r02j_prepass_ab.jsonl; tail -3 $OUT/r02j_prepass_ab.err
