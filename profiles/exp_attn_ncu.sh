#!/bin/bash
# full ncu capture (source-level) of one attention launch of the second forward
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r01d}
ENC_REPS=2 ncu --set full --import-source on --clock-control none -k regex:attention_tc -s 13 -c 1 -f -o $OUT/${TAG}_attention \
    python profiles/encoder_once.py > $OUT/${TAG}_attn_ncu.log 2>&1
tail -2 $OUT/${TAG}_attn_ncu.log
