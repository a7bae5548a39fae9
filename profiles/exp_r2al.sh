#!/bin/bash
# r02al: smoke() with the batch-1 forward added (few-token kernels in the driver's kernel list)
OUT=gpurun_out; mkdir -p $OUT
( time python __graft_entry__.py smoke ) > $OUT/r02al_smoke.log 2>&1; tail -n 6 $OUT/r02al_smoke.log
