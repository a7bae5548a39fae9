#!/bin/bash
# r02an: services / store GPU tests after the _Column._pending tidy-up
OUT=gpurun_out; mkdir -p $OUT
( time timeout 80 python -m pytest tests/test_services_gpu.py -x -q -m gpu ) > $OUT/r02an_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02an_pytest.log | tail -n 4 | cut -c1-200
