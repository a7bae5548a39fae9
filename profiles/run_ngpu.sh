#!/bin/bash
# N-GPU pass under `gpurun --gpus N`: shard parity test at world N, the bench at N (fused peer-store exchange),
# and the HBM-bound B=128 point.   bash profiles/run_ngpu.sh <tag> <N>
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r01c}; N=${2:-8}
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_${N}gpu_smi.txt
nvidia-smi topo -m > $OUT/${TAG}_${N}gpu_topo.txt 2>&1
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -x -q > $OUT/${TAG}_${N}gpu_pytest.log 2>&1
tail -n 5 $OUT/${TAG}_${N}gpu_pytest.log
: > $OUT/${TAG}_${N}gpu_bench.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 >> $OUT/${TAG}_${N}gpu_bench.jsonl 2>> $OUT/${TAG}_${N}gpu_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --batch 128 --no-encoder --no-cpu-baseline >> $OUT/${TAG}_${N}gpu_bench.jsonl 2>> $OUT/${TAG}_${N}gpu_bench.err
cat $OUT/${TAG}_${N}gpu_bench.jsonl; tail -n 5 $OUT/${TAG}_${N}gpu_bench.err
