#!/bin/bash
# r02ai: few-token path with 256-element K ranges per warp (12 warps per CTA): parity + latency
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_ner_gpu.py -x -q -m gpu ) > $OUT/r02ai_pytest_enc.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02ai_pytest_enc.log | tail -n 6 | cut -c1-300
sed -i 's/for S in (8, 12, 24, 48, 64, 65):/for S in (8, 12, 24, 32, 48, 64, 65):/; s/N.tune(enc_skinny=mode)/N.tune(enc_skinny=2 * mode)/' profiles/enc_latency.py
timeout 300 python profiles/enc_latency.py > $OUT/r02ai_enc_latency.jsonl 2> $OUT/r02ai_enc_latency.err
cat $OUT/r02ai_enc_latency.jsonl; tail -3 $OUT/r02ai_enc_latency.err
