#!/bin/bash
# r02f: full GPU suite (1 GPU) after the configs[0] test fix
OUT=gpurun_out; mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -x -q -s ) > $OUT/r02f_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02f_pytest.log | tail -n 12
cat $OUT/r02_config0.json 2>/dev/null | head -80
