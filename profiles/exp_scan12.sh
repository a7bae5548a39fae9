#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_services_gpu.py -m gpu -x -q ) > $OUT/kbs_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/kbs_pytest.log
: > $OUT/r01f_sweep_10M.jsonl
for B in 1 8 32 128 256 1024 2048 4096; do
  timeout 300 python bench.py --rows 10000000 --batch $B --steps 8 --warmup 3 --no-encoder --no-cpu-baseline >> $OUT/r01f_sweep_10M.jsonl 2>> $OUT/r01f_sweep_10M.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r01f_sweep_10M.jsonl"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], round(d["value"]), round(d["e2e"]["value"]), r["bound"], round(r["frac"], 3), round(r["kernel_us"]), d["clocks"]["sm_mhz"], d["ids_match_host_device"])
PY
tail -3 $OUT/r01f_sweep_10M.err
