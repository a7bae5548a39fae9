#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_scan_gpu.py -m gpu -x -q ) > $OUT/kbs_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/kbs_pytest.log
: > $OUT/kbs2_sweep.jsonl
for B in 1024 128; do for V in scan_kbs=3 scan_kbs=4 scan_kbs=6 scan_kbs=3; do
  timeout 300 python bench.py --rows 10000000 --batch $B --steps 8 --warmup 3 --no-encoder --no-cpu-baseline --tune $V >> $OUT/kbs2_sweep.jsonl 2>> $OUT/kbs2_sweep.err
done; done
python - <<'PY'
import json
for l in open("gpurun_out/kbs2_sweep.jsonl"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], d["config"].get("tune"), round(d["value"]), r["bound"], round(r["frac"], 3), round(r["kernel_us"]), d["clocks"]["sm_mhz"], d["ids_match_host_device"])
PY
tail -3 $OUT/kbs2_sweep.err
