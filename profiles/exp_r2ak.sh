#!/bin/bash
# r02ak: final state of the round: full GPU suite, smoke, bench + reference arm
OUT=gpurun_out; mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > $OUT/r02ak_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02ak_pytest.log | tail -n 6
( time python __graft_entry__.py smoke ) > $OUT/r02ak_smoke.log 2>&1; tail -n 2 $OUT/r02ak_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/r02ak_bench.json 2> $OUT/r02ak_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02ak_bench.json'))
print(json.dumps(d.get('config1_point'))[:900]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['hbm_point']['roofline']['frac'], d['ridge_point']['roofline']['frac'], d['encoder']['value'], d['encoder']['e2e_text']['value'], d['checks'])
PY
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $OUT/r02ak_bench_ref.json 2> $OUT/r02ak_bench_ref.err
cut -c1-400 $OUT/r02ak_bench_ref.json
