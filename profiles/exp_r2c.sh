#!/bin/bash
# r02c: encoder tests (feeder pipeline, outlier-channel stress), services tests, then the full default bench line
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_encoder_gpu.py tests/test_services_gpu.py tests/test_ner_gpu.py -m gpu -x -q ) > $OUT/r02c_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 $OUT/r02c_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/r02c_bench.json 2> $OUT/r02c_bench.err
cat $OUT/r02c_bench.json; tail -5 $OUT/r02c_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $OUT/r02c_bench_ref.json 2> $OUT/r02c_bench_ref.err
cat $OUT/r02c_bench_ref.json; tail -5 $OUT/r02c_bench_ref.err
