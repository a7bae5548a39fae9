#!/bin/bash
# r02ag: 2-GPU shard tests after making the sharded-build worker compare batched and single searches on the same vectors
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_shard_gpu.py tests/test_bench_gpu.py -m gpu -x -q ) > $OUT/r02ag_2gpu_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02ag_2gpu_pytest.log | tail -n 12 | cut -c1-600
