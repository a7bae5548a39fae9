#!/bin/bash
# r02m: configs[1] latency probe (C ABI vs service layers), feeder timers, bench line
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python profiles/config1_probe.py > $OUT/r02m_config1_probe.txt 2>&1; cat $OUT/r02m_config1_probe.txt | grep -v INFO
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/r02m_bench.json 2> $OUT/r02m_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench.json'))
print(d['value'], d['roofline']['frac'], d['hbm_point']['roofline']['frac'], d['ridge_point']['roofline']['frac'], d['encoder']['value'])
print(json.dumps(d['encoder']['e2e_text'])[:1500])
PY
