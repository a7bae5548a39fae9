#!/bin/bash
# r02h: ncu evidence for the kernels as shipped (scan at every per-GPU shard size of the 1/2/4/8-GPU runs, encoder layer),
# launch list of the default bench command.  Captures are summarised ON the box (gpurun copies back <= 64 MiB).
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max.per_second"
# can the reference's own engines run on this box?
( python -c "import sentence_transformers" 2>&1 | tail -n 1; python -c "import pymilvus" 2>&1 | tail -n 1; python -c "import milvus_lite" 2>&1 | tail -n 1; ls -la ~/.cache/huggingface 2>&1 | head -n 5; pip download sentence-transformers --no-deps -d /tmp/x 2>&1 | tail -n 1 ) > $OUT/r02h_probe.txt 2>&1
cat $OUT/r02h_probe.txt
# full capture at the 8-GPU shard size; DRAM bytes / time / pipe utilisation (2 passes) at the 1-, 2-, 4-GPU shard sizes
ROWS=12500000 BATCHES=1024,256,128 timeout 900 $NCU --profile-from-start off --set full --import-source on -k regex:scan_tc -f -o $OUT/r02h_scan_12500000 \
    python profiles/ncu_scan.py > $OUT/r02h_ncu_scan_12500000.log 2>&1
for cfg in "100000000 1024,256,128" "50000000 1024,256,128" "25000000 1024,256,128"; do
  set -- $cfg
  ROWS=$1 BATCHES=$2 timeout 900 $NCU --profile-from-start off --metrics $M -k regex:scan_tc --csv --log-file $OUT/r02h_scan_$1_metrics.csv \
      python profiles/ncu_scan.py > $OUT/r02h_ncu_scan_$1.log 2>&1
  tail -n 1 $OUT/r02h_ncu_scan_$1.log
done
# encoder: the five kernels of layer 1 of the second forward, full; then DRAM bytes + time of all 63 launches of one forward
ENC_REPS=2 timeout 900 $NCU --set full --import-source on -k regex:'gemm_tc|attention' -s 65 -c 5 -f -o $OUT/r02h_encoder_layer \
    python profiles/encoder_once.py > $OUT/r02h_ncu_encoder.log 2>&1
ENC_REPS=2 timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 63 -c 63 --csv --log-file $OUT/r02h_encoder_launches.csv python profiles/encoder_once.py > /dev/null 2>&1
# launch list of the default bench command (first 600 launches of our kernels)
timeout 1200 $NCU --metrics gpu__time_duration.sum -k regex:'scan_|merge_kernel|finalise_kernel|gemm_tc|attention|layernorm|embed_ln|pool_normalise|bf16|token_head' -c 600 --csv --log-file $OUT/r02h_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02h_launches_bench.log 2>&1
python profiles/summarise.py $OUT/r02h_scan_12500000.ncu-rep $OUT/r02h_encoder_layer.ncu-rep > $OUT/r02h_ncu_summary.txt 2>&1
ncu -i $OUT/r02h_scan_12500000.ncu-rep --page source --csv > $OUT/r02h_src_scan.csv 2>/dev/null
python profiles/stalls.py $OUT/r02h_src_scan.csv 25 > $OUT/r02h_stalls_scan_b1024.txt 2>&1; rm -f $OUT/r02h_src_scan.csv
du -sh $OUT; ls -la $OUT | grep r02h
