#!/bin/bash
# r02v: (second try: queue in local memory, one drain call per tile) queued admissions in the scan epilogue (+ pre-pass without list storage, kbs = 4 issue loop, block-per-query bound
# kernel): parity, small-table A/B, and same-box A/B against the r02s library on 10 M rows under sustained load
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests/test_scan_gpu.py -x -q -m gpu ) > $OUT/r02v_pytest_scan.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02v_pytest_scan.log | tail -n 5
timeout 600 python profiles/small_table_ab.py 40474 > $OUT/r02v_small_table_ab.jsonl 2> $OUT/r02v_small_table_ab.err
cut -c1-330 $OUT/r02v_small_table_ab.jsonl
OLD=rag-project-icd10_b200/csrc/build/ab/libicdrag_r02s.so
: > $OUT/r02v_big_ab.jsonl
for spec in "128 400" "256 300" "1024 100" "32 500"; do
  set -- $spec
  for lib in old new old new; do
    if [ $lib = old ]; then export PROF_LIB=$OLD; else unset PROF_LIB; fi
    ROWS=10000000 BATCH=$1 STEPS=$2 WARM=20 VARIANTS="scan_pair=-1" timeout 300 python profiles/scan_ab.py 2>/dev/null | sed "s/^{/{\"lib\": \"$lib\", /" >> $OUT/r02v_big_ab.jsonl
  done
done
unset PROF_LIB
cut -c1-200 $OUT/r02v_big_ab.jsonl
