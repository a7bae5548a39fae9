#!/bin/bash
# r02ac: few-token path: activation rows prefetched into L1 in one sweep, two MMA chains per tile
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_ner_gpu.py -x -q -m gpu ) > $OUT/r02ac_pytest_enc.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02ac_pytest_enc.log | tail -n 8
timeout 300 python profiles/enc_latency.py > $OUT/r02ac_enc_latency.jsonl 2> $OUT/r02ac_enc_latency.err
cat $OUT/r02ac_enc_latency.jsonl; tail -3 $OUT/r02ac_enc_latency.err
S=12 ncu --clock-control none --metrics gpu__time_duration.sum -s 126 -c 63 --csv --log-file $OUT/r02ac_b1_launches_skinny.csv python profiles/enc_once_b1.py > /dev/null 2>&1
S=12 SKINNY=0 ncu --clock-control none --metrics gpu__time_duration.sum -s 126 -c 63 --csv --log-file $OUT/r02ac_b1_launches_tile.csv python profiles/enc_once_b1.py > /dev/null 2>&1
