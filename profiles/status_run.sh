#!/bin/bash
# Status pass under gpurun: GPU tests, smoke, the default bench line (+ reference arm), encoder launch list.
#   bash profiles/status_run.sh <tag> [sweep]
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r01c}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -n 5 $OUT/${TAG}_pytest.log
( time python __graft_entry__.py smoke ) > $OUT/${TAG}_smoke.log 2>&1
tail -n 3 $OUT/${TAG}_smoke.log
( time python bench.py ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cat $OUT/${TAG}_bench_ref.json
ENC_REPS=2 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 63 -c 63 --csv \
    --log-file $OUT/${TAG}_encoder_launches.csv python profiles/encoder_once.py > /dev/null 2>&1
if [ "${2:-}" = "sweep" ]; then
for B in 1 8 32 128 256 1024 4096; do
  python bench.py --rows 10000000 --batch $B --steps 5 --warmup 3 --no-encoder --no-cpu-baseline >> $OUT/${TAG}_sweep_10M.jsonl 2>> $OUT/${TAG}_sweep_10M.err
done
fi
