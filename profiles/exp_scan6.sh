#!/bin/bash
# CTA-pair scan: parity tests (bounded by timeout: a pipeline bug would hang), then 10 M rows, pair/split variants
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_scan_gpu.py -m gpu -x -q ) > $OUT/pair_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 $OUT/pair_pytest.log
: > $OUT/pair_sweep.jsonl
for B in 1024 256 4096; do for V in scan_pair=0 scan_pair=-1 scan_pair=-1,scan_qtmem=8 scan_pair=-1,scan_qtmem=4; do
  timeout 300 python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --no-encoder --no-cpu-baseline --tune $V >> $OUT/pair_sweep.jsonl 2>> $OUT/pair_sweep.err
done; done
python - <<'PY'
import json
for l in open("gpurun_out/pair_sweep.jsonl"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], d["config"].get("tune"), round(d["value"]), r["bound"], round(r["frac"], 3), round(r["kernel_us"]), d["ids_match_host_device"])
PY
tail -5 $OUT/pair_sweep.err
