#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/tmax16_sweep.jsonl
for B in 2048 4096; do for V in scan_tmax=8 scan_tmax=16 scan_tmax=8 scan_tmax=16; do
  timeout 300 python bench.py --rows 10000000 --batch $B --steps 8 --warmup 3 --no-encoder --no-cpu-baseline --tune $V >> $OUT/tmax16_sweep.jsonl 2>> $OUT/tmax16_sweep.err
done; done
python - <<'PY'
import json
for l in open("gpurun_out/tmax16_sweep.jsonl"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], d["config"].get("tune"), round(d["value"]), round(d["ms_per_step"],3), r["bound"], round(r["frac"], 3), round(r["kernel_us"]), d["launches_per_step"], d["clocks"]["sm_mhz"], d["ids_match_host_device"])
PY
tail -3 $OUT/tmax16_sweep.err
