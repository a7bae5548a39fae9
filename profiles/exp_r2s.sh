#!/bin/bash
# r02s: small-table regime of the tensor scan (pre-pass below 512 k rows, counted inserts, 128-set bound): parity + A/B
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests/test_scan_gpu.py -x -q -m gpu ) > $OUT/r02s_pytest_scan.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02s_pytest_scan.log | tail -n 8
timeout 600 python profiles/small_table_ab.py 40474 2048 300000 > $OUT/r02s_small_table_ab.jsonl 2> $OUT/r02s_small_table_ab.err
cat $OUT/r02s_small_table_ab.jsonl | cut -c1-420
