#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/kbs_sweep.jsonl
for B in 8 32; do for V in scan_qsplit=1 scan_qsplit=1,scan_kbs=3; do
  timeout 300 python bench.py --rows 10000000 --batch $B --steps 10 --warmup 3 --no-encoder --no-cpu-baseline --tune $V >> $OUT/kbs_sweep.jsonl 2>> $OUT/kbs_sweep.err
done; done
for B in 256 1024; do for V in scan_kbs=2 scan_kbs=3 scan_kbs=2 scan_kbs=3; do
  timeout 300 python bench.py --rows 10000000 --batch $B --steps 8 --warmup 3 --no-encoder --no-cpu-baseline --tune $V >> $OUT/kbs_sweep.jsonl 2>> $OUT/kbs_sweep.err
done; done
python - <<'PY'
import json
for l in open("gpurun_out/kbs_sweep.jsonl"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], d["config"].get("tune"), round(d["value"]), r["bound"], round(r["frac"], 3), round(r["kernel_us"]), d["clocks"]["sm_mhz"], d["ids_match_host_device"])
PY
tail -3 $OUT/kbs_sweep.err
