#!/bin/bash
# r01d profiling pass (one B200): full captures of the kernels as they are now.
set -u
TAG=${1:-r01d}
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
# scan, CTA-pair mode, 10 M rows: tensor-bound (B=1024) and near the ridge (B=256); main-scan launch = 2nd scan_tc launch
for B in 1024 256; do
$NCU --set full --import-source on -k regex:scan_tc -s 3 -c 1 -f -o $OUT/${TAG}_scan_tc_b$B \
    python bench.py --rows 10000000 --batch $B --steps 1 --warmup 1 --no-encoder --no-cpu-baseline > $OUT/${TAG}_ncu_b$B.log 2>&1
done
# encoder: the five kernels of layer 1 of the second forward (63 launches per forward: embed, 12 x 5, LN, pool)
ENC_REPS=2 $NCU --set full --import-source on -k regex:'gemm_tc|attention' -s 65 -c 5 -f -o $OUT/${TAG}_encoder_layer \
    python profiles/encoder_once.py > $OUT/${TAG}_ncu_encoder.log 2>&1
ENC_REPS=2 $NCU --metrics gpu__time_duration.sum -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 63 -c 63 --csv \
    --log-file $OUT/${TAG}_encoder_launches.csv python profiles/encoder_once.py > /dev/null 2>&1
# launch list of the default bench command (our kernels only)
$NCU --metrics gpu__time_duration.sum -k regex:'scan_|merge_kernel|finalise_kernel|gemm_tc|attention|layernorm|embed_ln|pool_normalise|bf16' \
    --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
ls -la $OUT | grep $TAG
