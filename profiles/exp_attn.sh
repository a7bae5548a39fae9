#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_services_gpu.py -m gpu -x -q ) > $OUT/ln_pytest.log 2>&1
tail -n 3 $OUT/ln_pytest.log
for D in 0 1; do
ICD_ATTN_DBG=$D ENC_REPS=2 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'attention|layernorm' -s 13 -c 13 --csv \
    --log-file $OUT/attn_dbg$D.csv python profiles/encoder_once.py > /dev/null 2>&1
echo "dbg=$D"; python profiles/launch_summary.py $OUT/attn_dbg$D.csv | tail -2
done
for F in 1 1; do ICD_ENC_FUSED_LN=$F python profiles/encoder_time.py 2>&1 | tail -1; done | tee $OUT/ln_time.txt
