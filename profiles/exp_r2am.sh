#!/bin/bash
# r02am: ncu launch list of the final kernels (bench command at the 8-GPU shard size, scan part + config1_point)
OUT=gpurun_out; mkdir -p $OUT
OURS='regex:scan_|merge_kernel|finalise_kernel|bound_from|bf16'
timeout 150 ncu --clock-control none --metrics gpu__time_duration.sum -k "$OURS" --csv --log-file $OUT/r02am_launches_bench.csv \
    python bench.py --rows 12500000 --steps 2 --warmup 1 --no-cpu-baseline --no-encoder > $OUT/r02am_launches_bench.log 2>&1
echo "rc=$?"; wc -l $OUT/r02am_launches_bench.csv
