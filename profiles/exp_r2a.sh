#!/bin/bash
# r02a: where does the ridge (B = 256, 10 M rows) lose its time?  Same-box knob A/B + one full ncu capture.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv > $OUT/r02a_gpu.txt
ROWS=10000000 BATCH=256 STEPS=20 VARIANTS="scan_pair=-1;scan_kbs_pair=4;scan_kbs_pair=3;scan_kbs_pair=2;scan_sample=0;scan_qsplit=0;scan_pair=0;scan_pair=0,scan_kbs=6;scan_pair=0,scan_drift=0" \
  timeout 600 python profiles/scan_ab.py > $OUT/r02a_ab_b256.jsonl 2> $OUT/r02a_ab_b256.err
cat $OUT/r02a_ab_b256.jsonl
ROWS=10000000 BATCH=128 STEPS=20 VARIANTS="scan_pair=-1;scan_kbs=2;scan_kbs=6;scan_sample=0" \
  timeout 600 python profiles/scan_ab.py > $OUT/r02a_ab_b128.jsonl 2> $OUT/r02a_ab_b128.err
cat $OUT/r02a_ab_b128.jsonl
ROWS=12500000 BATCH=1024 STEPS=10 VARIANTS="scan_pair=-1;scan_sample=0;scan_drift=8;scan_drift=2" \
  timeout 600 python profiles/scan_ab.py > $OUT/r02a_ab_b1024.jsonl 2> $OUT/r02a_ab_b1024.err
cat $OUT/r02a_ab_b1024.jsonl
timeout 900 ncu --clock-control none --set full --import-source on -k regex:scan_tc -s 3 -c 1 -f -o $OUT/r02a_scan_tc_b256 \
    python bench.py --rows 10000000 --batch 256 --steps 1 --warmup 1 --no-encoder --no-cpu-baseline > $OUT/r02a_ncu_b256.log 2>&1
ls -la $OUT | grep r02a
