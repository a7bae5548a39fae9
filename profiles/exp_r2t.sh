#!/bin/bash
# r02t: ncu of the small-table scan (40 474 rows): B = 8192 (4 pre-pass + 4 main launches per search) and B = 1024
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
B=8192 $NCU --set full --import-source on -k regex:'scan_tc|bound_from|merge_kernel|finalise' -s 22 -c 11 -f -o $OUT/r02t_small_b8192 python profiles/small_table_once.py > $OUT/r02t_ncu_b8192.log 2>&1
B=1024 $NCU --set full --import-source on -k regex:'scan_tc|bound_from|merge_kernel|finalise' -s 10 -c 5 -f -o $OUT/r02t_small_b1024 python profiles/small_table_once.py > $OUT/r02t_ncu_b1024.log 2>&1
B=64 $NCU --set full --import-source on -k regex:'scan_tc|bound_from|merge_kernel|finalise' -s 10 -c 5 -f -o $OUT/r02t_small_b64 python profiles/small_table_once.py > $OUT/r02t_ncu_b64.log 2>&1
ls -la $OUT/r02t*
