#!/bin/bash
# r02y: few-token weight-streaming path of the encoder (skinny_linear.cu): parity + batch-1 latency A/B
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_ner_gpu.py -x -q -m gpu ) > $OUT/r02y_pytest_enc.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02y_pytest_enc.log | tail -n 25
timeout 300 python profiles/enc_latency.py > $OUT/r02y_enc_latency.jsonl 2> $OUT/r02y_enc_latency.err
cat $OUT/r02y_enc_latency.jsonl; tail -3 $OUT/r02y_enc_latency.err
