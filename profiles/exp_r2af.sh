#!/bin/bash
# r02af: 2-GPU pass of the final state: the three tests that need >= 2 GPUs, then the bench at N = 2
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/r02af_2gpu_smi.txt
( time timeout 900 python -m pytest tests/test_shard_gpu.py tests/test_bench_gpu.py -m gpu -x -q ) > $OUT/r02af_2gpu_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "INFO\|WARNING\|^$" $OUT/r02af_2gpu_pytest.log | tail -n 5
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $OUT/r02af_2gpu_bench.json 2> $OUT/r02af_2gpu_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02af_2gpu_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['hbm_point']['roofline']['frac'], d['ridge_point']['roofline']['frac'], d['encoder']['value'], d['encoder']['e2e_text']['value'], d['checks'])
PY
tail -3 $OUT/r02af_2gpu_bench.err
