#!/bin/bash
# r02i: does a smaller encoder batch (activations closer to the 126 MB L2) buy throughput under the power cap?
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02i_encoder_batch_sweep.txt
for B in 4096 2048 1024 512 256 4096; do
  ENC_B=$B ENC_S=64 ENC_REPS=$((81920 / B)) timeout 600 python profiles/encoder_time.py >> $OUT/r02i_encoder_batch_sweep.txt 2>> $OUT/r02i_encoder_batch_sweep.err
done
cat $OUT/r02i_encoder_batch_sweep.txt; tail -3 $OUT/r02i_encoder_batch_sweep.err
