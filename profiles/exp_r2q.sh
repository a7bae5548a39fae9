#!/bin/bash
# r02q: split-KV tensor-core attention for 128 < S <= 512: parity (encoder + NER suites)
OUT=gpurun_out; mkdir -p $OUT
( timeout 1200 python -m pytest tests/test_encoder_gpu.py tests/test_ner_gpu.py -m gpu -x -q ) > $OUT/r02q_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02q_pytest.log | tail -n 30
