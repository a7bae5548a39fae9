#!/bin/bash
# r02l: full GPU suite + default bench line with the slot-maxima pre-pass; encoder batch-size sweep
OUT=gpurun_out; mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > $OUT/r02l_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02l_pytest.log | tail -n 8
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/r02l_bench.json 2> $OUT/r02l_bench.err
cat $OUT/r02l_bench.json | cut -c1-3000
bash profiles/exp_r2i.sh
