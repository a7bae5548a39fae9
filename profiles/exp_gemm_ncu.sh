#!/bin/bash
# Full ncu capture (with source) of the four GEMMs of one encoder layer.
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-gemm}
ENC_REPS=2 timeout 900 ncu --clock-control none --set full --import-source on -k regex:gemm_tc -s 52 -c 4 -f -o $OUT/${TAG}_gemms python profiles/encoder_once.py > $OUT/${TAG}_ncu.log 2>&1
tail -n 3 $OUT/${TAG}_ncu.log
