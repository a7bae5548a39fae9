#!/bin/bash
# 2-GPU pass under `gpurun --gpus 2`: shard parity test, the bench at N=2 with both exchanges, the reference arm.
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r01b}
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_2gpu_smi.txt
nvidia-smi topo -m > $OUT/${TAG}_2gpu_topo.txt 2>&1
timeout 900 python -m pytest tests/test_shard_gpu.py -m gpu -x -q > $OUT/${TAG}_2gpu_pytest.log 2>&1
tail -n 5 $OUT/${TAG}_2gpu_pytest.log
: > $OUT/${TAG}_2gpu_bench.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 >> $OUT/${TAG}_2gpu_bench.jsonl 2>> $OUT/${TAG}_2gpu_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --exchange 0 --no-encoder >> $OUT/${TAG}_2gpu_bench.jsonl 2>> $OUT/${TAG}_2gpu_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --batch 128 --no-encoder >> $OUT/${TAG}_2gpu_bench.jsonl 2>> $OUT/${TAG}_2gpu_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 >> $OUT/${TAG}_2gpu_bench.jsonl 2>> $OUT/${TAG}_2gpu_bench.err
cat $OUT/${TAG}_2gpu_bench.jsonl; tail -n 5 $OUT/${TAG}_2gpu_bench.err
