#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for D in 0 1 2 3; do
ICD_ATTN_DBG=$D ENC_REPS=2 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'attention|layernorm' -s 13 -c 13 --csv \
    --log-file $OUT/attn_dbg$D.csv python profiles/encoder_once.py > /dev/null 2>&1
echo "dbg=$D"; python profiles/launch_summary.py $OUT/attn_dbg$D.csv | tail -2
done
