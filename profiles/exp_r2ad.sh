#!/bin/bash
# r02ad: few-token attention kernel + pipelined host-buffer searches: parity, latency, config1 through C ABI / service
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_ner_gpu.py tests/test_scan_gpu.py -x -q -m gpu -k "not config3" ) > $OUT/r02ad_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02ad_pytest.log | tail -n 8
timeout 300 python profiles/enc_latency.py > $OUT/r02ad_enc_latency.jsonl 2> $OUT/r02ad_enc_latency.err
cat $OUT/r02ad_enc_latency.jsonl; tail -3 $OUT/r02ad_enc_latency.err
timeout 300 python profiles/small_table_ab.py 40474 > $OUT/r02ad_small_table_ab.jsonl 2> $OUT/r02ad_small_table_ab.err
grep '"keep_f32": true' $OUT/r02ad_small_table_ab.jsonl | cut -c1-330 | tail -3
