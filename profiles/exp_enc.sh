#!/bin/bash
# Encoder experiment pass: parity tests, throughput at a few shapes, per-kernel times of one layer.
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-enc}
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
tail -n 5 $OUT/${TAG}_pytest.log
timeout 600 python - > $OUT/${TAG}_bench.txt 2>&1 <<'PY'
import importlib, json, os, sys, torch
sys.path.insert(0, os.getcwd())
E = importlib.import_module("rag-project-icd10_b200.engine.encoder")
peaks = {"bf16_tflops_sustained": 1399.4}
dev = torch.device("cuda", 0)
for S, B in ((64, 4096), (128, 2048), (24, 4096)):
    print(S, B, json.dumps(E.bench_encoder(dev, peaks, batch=B, seq=S, steps=5, warmup=3)))
PY
cat $OUT/${TAG}_bench.txt
ENC_REPS=2 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 86 -c 8 --csv --log-file $OUT/${TAG}_layer_times.csv python profiles/encoder_once.py > /dev/null 2>&1
cut -d, -f5,15 $OUT/${TAG}_layer_times.csv | tail -n 9
