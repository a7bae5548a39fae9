#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/b128_sweep.jsonl
for V in scan_kbs=2 scan_qsplit=1 scan_kbs=3 scan_kbs=1 scan_kbs=2 scan_kbs=3,scan_qsplit=1; do
  timeout 300 python bench.py --rows 10000000 --batch 128 --steps 10 --warmup 3 --no-encoder --no-cpu-baseline --tune $V >> $OUT/b128_sweep.jsonl 2>> $OUT/b128_sweep.err
done
for V in scan_kbs=2 scan_kbs=3; do
  timeout 300 python bench.py --rows 10000000 --batch 32 --steps 10 --warmup 3 --no-encoder --no-cpu-baseline --tune $V >> $OUT/b128_sweep.jsonl 2>> $OUT/b128_sweep.err
done
python - <<'PY'
import json
for l in open("gpurun_out/b128_sweep.jsonl"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], d["config"].get("tune"), round(d["value"]), r["bound"], round(r["frac"], 3), round(r["kernel_us"]), d["clocks"]["sm_mhz"], d["ids_match_host_device"])
PY
tail -3 $OUT/b128_sweep.err
