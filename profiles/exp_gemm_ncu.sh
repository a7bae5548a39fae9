#!/bin/bash
# full ncu capture (source-level) of the four GEMMs of encoder layer 1, second forward
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r01d}
ENC_REPS=2 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 52 -c 4 -f -o $OUT/${TAG}_gemm_layer \
    python profiles/encoder_once.py > $OUT/${TAG}_gemm_ncu.log 2>&1
tail -3 $OUT/${TAG}_gemm_ncu.log; ls -la $OUT/*.ncu-rep
