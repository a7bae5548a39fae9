#!/bin/bash
# r02ah: ncu of the few-token linear kernels at 12 and 32 tokens (what does a second 16-token tile cost?)
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
S=12 $NCU --set full --import-source on -k regex:skinny_linear -s 96 -c 4 -f -o $OUT/r02ah_skinny_s12 python profiles/enc_once_b1.py > $OUT/r02ah_ncu_s12.log 2>&1
S=32 $NCU --set full --import-source on -k regex:skinny_linear -s 96 -c 4 -f -o $OUT/r02ah_skinny_s32 python profiles/enc_once_b1.py > $OUT/r02ah_ncu_s32.log 2>&1
ls -la $OUT/r02ah*
