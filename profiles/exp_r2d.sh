#!/bin/bash
# r02d: configs[0] at full size (build + query + batch-1 latency), configs[3] independent checker, full GPU suite
OUT=gpurun_out; mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -x -q -s ) > $OUT/r02d_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02d_pytest.log | tail -n 25
cat $OUT/r02_config0.json 2>/dev/null | head -60
( time python __graft_entry__.py smoke ) > $OUT/r02d_smoke.log 2>&1; tail -n 3 $OUT/r02d_smoke.log
