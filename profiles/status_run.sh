#!/bin/bash
# Status pass under gpurun: GPU tests, smoke, the default bench line, a batch sweep on 10 M rows.
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r01b}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -n 5 $OUT/${TAG}_pytest.log
( time python __graft_entry__.py smoke ) > $OUT/${TAG}_smoke.log 2>&1
tail -n 3 $OUT/${TAG}_smoke.log
( time python bench.py ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cat $OUT/${TAG}_bench_ref.json
for B in 1 8 32 128 256 1024 4096; do
  python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --no-encoder --no-cpu-baseline >> $OUT/${TAG}_sweep_10M.jsonl 2>> $OUT/${TAG}_sweep_10M.err
done
python - <<'PY'
import json
for l in open("gpurun_out/%s_sweep_10M.jsonl" % "r01b"):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["batch"], round(d["value"]), round(d["e2e"]["value"]), r["bound"], round(r["frac"], 3), round(r["kernel_us"]), d["gpu_launches"])
PY
