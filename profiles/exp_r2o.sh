#!/bin/bash
# r02o: programmatic dependent launch across the encoder's kernels: parity with it on, batch-1 latency and throughput A/B
OUT=gpurun_out; mkdir -p $OUT
( ICD_TEST_ENC_PDL=1 timeout 1200 python -m pytest tests/test_encoder_gpu.py tests/test_ner_gpu.py -m gpu -x -q ) > $OUT/r02o_pytest_pdl.log 2>&1
echo "pytest(pdl) rc=$?"; tail -n 3 $OUT/r02o_pytest_pdl.log
for P in 0 1 0 1; do ENC_PDL=$P timeout 300 python profiles/latency_probe.py 2>/dev/null | grep -v INFO >> $OUT/r02o_latency.txt; done
cat $OUT/r02o_latency.txt
for P in 0 1 0 1; do ENC_PDL=$P ENC_REPS=40 timeout 300 python profiles/encoder_time.py 2>/dev/null | tail -n 1 >> $OUT/r02o_throughput.txt; done
cat $OUT/r02o_throughput.txt
