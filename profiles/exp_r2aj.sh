#!/bin/bash
# r02aj: few-token path: shape by token count (<= 16: 768-element K ranges, else 256), second pass on other CTAs (blockIdx.y)
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_ner_gpu.py -x -q -m gpu ) > $OUT/r02aj_pytest_enc.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02aj_pytest_enc.log | tail -n 6 | cut -c1-300
timeout 300 python profiles/enc_latency.py > $OUT/r02aj_enc_latency.jsonl 2> $OUT/r02aj_enc_latency.err
cat $OUT/r02aj_enc_latency.jsonl; tail -3 $OUT/r02aj_enc_latency.err
