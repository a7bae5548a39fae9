#!/bin/bash
# r02h: ncu evidence for the kernels as shipped (scan at every per-GPU shard size of the 1/2/4/8-GPU runs, encoder layer),
# launch list of the default bench command
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
# can the reference's own engines run on this box?
( python -c "import sentence_transformers" 2>&1 | tail -n 1; python -c "import pymilvus" 2>&1 | tail -n 1; python -c "import milvus_lite" 2>&1 | tail -n 1; ls -la ~/.cache/huggingface 2>&1 | head -n 5; pip download sentence-transformers --no-deps -d /tmp/x 2>&1 | tail -n 1 ) > $OUT/r02h_probe.txt 2>&1
cat $OUT/r02h_probe.txt
for cfg in "100000000 1024,256,128" "50000000 1024" "25000000 1024" "12500000 1024,256,128"; do
  set -- $cfg
  ROWS=$1 BATCHES=$2 timeout 900 $NCU --profile-from-start off --set full --import-source on -k regex:scan_tc -f -o $OUT/r02h_scan_$1 \
      python profiles/ncu_scan.py > $OUT/r02h_ncu_scan_$1.log 2>&1
  tail -n 2 $OUT/r02h_ncu_scan_$1.log
done
# encoder: the five kernels of layer 1 of the second forward, full; then DRAM bytes + time of all 63 launches of one forward
ENC_REPS=2 timeout 900 $NCU --set full --import-source on -k regex:'gemm_tc|attention' -s 65 -c 5 -f -o $OUT/r02h_encoder_layer \
    python profiles/encoder_once.py > $OUT/r02h_ncu_encoder.log 2>&1
ENC_REPS=2 timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 63 -c 63 --csv --log-file $OUT/r02h_encoder_launches.csv python profiles/encoder_once.py > /dev/null 2>&1
# launch list of the default bench command (first 600 launches of our kernels)
timeout 1200 $NCU --metrics gpu__time_duration.sum -k regex:'scan_|merge_kernel|finalise_kernel|gemm_tc|attention|layernorm|embed_ln|pool_normalise|bf16|token_head' -c 600 --csv --log-file $OUT/r02h_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02h_launches_bench.log 2>&1
ls -la $OUT | grep r02h
