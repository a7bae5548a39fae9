#!/bin/bash
# r02b: specialised (unrolled, ~3 instr / MMA) vs generic MMA issue loop of the tensor scan; parity first.
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_scan_gpu.py -m gpu -x -q ) > $OUT/r02b_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/r02b_pytest.log
ROWS=10000000 BATCH=256 STEPS=150 WARM=100 VARIANTS="scan_generic=1;scan_generic=0;scan_generic=0,scan_kbs_pair=3" \
  timeout 600 python profiles/scan_ab.py > $OUT/r02b_ab_b256.jsonl 2> $OUT/r02b_ab_b256.err
cat $OUT/r02b_ab_b256.jsonl
ROWS=10000000 BATCH=128 STEPS=200 WARM=150 VARIANTS="scan_generic=1;scan_generic=0;scan_generic=0,scan_kbs=2" \
  timeout 600 python profiles/scan_ab.py > $OUT/r02b_ab_b128.jsonl 2> $OUT/r02b_ab_b128.err
cat $OUT/r02b_ab_b128.jsonl
ROWS=12500000 BATCH=1024 STEPS=40 WARM=30 VARIANTS="scan_generic=1;scan_generic=0;scan_generic=0,scan_kbs_pair=3" \
  timeout 600 python profiles/scan_ab.py > $OUT/r02b_ab_b1024.jsonl 2> $OUT/r02b_ab_b1024.err
cat $OUT/r02b_ab_b1024.jsonl
ROWS=10000000 BATCH=4096 STEPS=12 WARM=8 VARIANTS="scan_generic=1;scan_generic=0" \
  timeout 600 python profiles/scan_ab.py > $OUT/r02b_ab_b4096.jsonl 2> $OUT/r02b_ab_b4096.err
cat $OUT/r02b_ab_b4096.jsonl
timeout 900 python bench.py --no-encoder --no-cpu-baseline --steps 20 --warmup 5 > $OUT/r02b_bench.json 2> $OUT/r02b_bench.err
cat $OUT/r02b_bench.json
timeout 900 ncu --clock-control none --set full --import-source on -k regex:scan_tc -s 3 -c 1 -f -o $OUT/r02b_scan_tc_b256 \
    python bench.py --rows 10000000 --batch 256 --steps 1 --warmup 1 --no-encoder --no-cpu-baseline --no-points > $OUT/r02b_ncu_b256.log 2>&1
ls -la $OUT | grep r02b
