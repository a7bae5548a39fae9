#!/bin/bash
# query-tile K split (two accumulator buffers) on/off: parity tests, then 10 M rows at the tensor-bound batch sizes
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_scan_gpu.py -m gpu -x -q ) > $OUT/qsplit_pytest.log 2>&1
tail -n 3 $OUT/qsplit_pytest.log
: > $OUT/qsplit_sweep.jsonl
for B in 256 1024 4096; do for Q in 0 1; do
  timeout 600 python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --no-encoder --no-cpu-baseline --tune scan_qsplit=$Q >> $OUT/qsplit_sweep.jsonl 2>> $OUT/qsplit_sweep.err
done; done
python - <<'PY'
import json
for l in open("gpurun_out/qsplit_sweep.jsonl"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], d["config"].get("tune"), round(d["value"]), r["bound"], round(r["frac"], 3), round(r["kernel_us"]))
PY
