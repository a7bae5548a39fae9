#!/bin/bash
# r02g (gpurun --gpus N): shard parity + sharded build at world N, then the bench at N with its start-up checks
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > $OUT/r02n_${N}gpu_smi.txt
( time timeout 1500 python -m pytest tests/test_shard_gpu.py tests/test_bench_gpu.py -m gpu -x -q ) > $OUT/r02n_${N}gpu_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 $OUT/r02n_${N}gpu_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r02n_${N}gpu_bench.json 2> $OUT/r02n_${N}gpu_bench.err
cat $OUT/r02n_${N}gpu_bench.json; tail -n 5 $OUT/r02n_${N}gpu_bench.err
