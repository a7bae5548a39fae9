#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_services_gpu.py -m gpu -x -q ) > $OUT/ln_pytest.log 2>&1
tail -n 15 $OUT/ln_pytest.log
for F in 1 0 1 0; do ICD_ENC_FUSED_LN=$F python profiles/encoder_time.py 2>&1 | tail -1; done | tee $OUT/ln_time.txt
ENC_REPS=2 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 64 -c 64 --csv \
    --log-file $OUT/ln_encoder_launches.csv python profiles/encoder_once.py > /dev/null 2>&1
python profiles/launch_summary.py $OUT/ln_encoder_launches.csv
