#!/bin/bash
# Profiling pass run under gpurun (one B200).  Writes everything to gpurun_out/; the summaries
# that are judged get copied into profiles/ by profiles/summarise.py on the CPU box.
#   bash profiles/run_profiles.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
OURS='regex:scan_|merge_kernel|finalise_kernel|gemm_tc|attention|layernorm|embed_ln|pool_normalise|bf16'

# (a) launch list of the bench command itself (our kernels only; torch's corpus generator is filtered out)
$NCU --metrics gpu__time_duration.sum -k "$OURS" --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1

# (b) full captures of the dominant kernels
#     the bench workload itself (100 M rows, B = 1024): second main-scan launch
$NCU --set full --import-source on -k regex:scan_tc -s 3 -c 1 -f -o $OUT/${TAG}_scan_tc_100M_b1024 \
    python bench.py --batch 1024 --steps 1 --warmup 1 --no-encoder --no-cpu-baseline > $OUT/${TAG}_ncu_100M_b1024.log 2>&1
#     10 M rows (BASELINE configs[3]) at the tensor-bound, ridge and HBM-bound batch sizes
for B in 4096 256 128; do
$NCU --set full --import-source on -k regex:scan_tc -s 3 -c 1 -f -o $OUT/${TAG}_scan_tc_b$B \
    python bench.py --rows 10000000 --batch $B --steps 1 --warmup 1 --no-encoder --no-cpu-baseline > $OUT/${TAG}_ncu_b$B.log 2>&1
done
$NCU --set full --import-source on -k regex:scan_stream -s 1 -c 1 -f -o $OUT/${TAG}_scan_stream_b1 \
    python bench.py --rows 10000000 --batch 1 --steps 1 --warmup 1 --no-encoder --no-cpu-baseline > $OUT/${TAG}_ncu_b1.log 2>&1
# encoder: one layer's worth of kernels of the second forward (1 + 12*7 + 1 launches per forward)
ENC_REPS=2 $NCU --set full --import-source on -k regex:'gemm_tc|attention|layernorm' -s 93 -c 7 -f -o $OUT/${TAG}_encoder_layer \
    python profiles/encoder_once.py > $OUT/${TAG}_ncu_encoder.log 2>&1
ENC_REPS=2 $NCU --metrics gpu__time_duration.sum -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 63 -c 63 --csv \
    --log-file $OUT/${TAG}_encoder_launches.csv python profiles/encoder_once.py > /dev/null 2>&1
ls -la $OUT
