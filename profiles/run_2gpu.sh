#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/2gpu_smi.txt
timeout 900 python -m pytest tests/test_shard_gpu.py -m gpu -x -q > $OUT/2gpu_pytest.log 2>&1
tail -n 5 $OUT/2gpu_pytest.log
for EX in 1 0; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --exchange $EX --no-encoder >> $OUT/2gpu_bench.jsonl 2>> $OUT/2gpu_bench.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 >> $OUT/2gpu_bench.jsonl 2>> $OUT/2gpu_bench.err
tail -n 3 $OUT/2gpu_bench.jsonl; tail -n 5 $OUT/2gpu_bench.err
