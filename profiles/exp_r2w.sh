#!/bin/bash
# r02w: prologue (query tile -> TMEM) with two K blocks per step; ncu of the small-table scan with the queued epilogue
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python profiles/small_table_ab.py 40474 > $OUT/r02w_small_table_ab.jsonl 2> $OUT/r02w_small_table_ab.err
cut -c1-330 $OUT/r02w_small_table_ab.jsonl | head -6
NCU="ncu --clock-control none"
B=8192 $NCU --set full --import-source on -k regex:'scan_tc' -s 16 -c 8 -f -o $OUT/r02w_small_b8192 python profiles/small_table_once.py > $OUT/r02w_ncu_b8192.log 2>&1
B=8192 $NCU --metrics gpu__time_duration.sum -k regex:'scan_tc|bound_from|merge_kernel|finalise|bf16' -s 0 -c 60 --csv --log-file $OUT/r02w_launches_b8192.csv python profiles/small_table_once.py > /dev/null 2>&1
B=64 $NCU --metrics gpu__time_duration.sum -k regex:'scan_tc|bound_from|merge_kernel|finalise|bf16' -s 0 -c 60 --csv --log-file $OUT/r02w_launches_b64.csv python profiles/small_table_once.py > /dev/null 2>&1
ls -la $OUT/r02w*
