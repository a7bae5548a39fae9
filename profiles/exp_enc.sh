#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -x -q > $OUT/enc_pytest.log 2>&1
tail -n 5 $OUT/enc_pytest.log
python - > $OUT/enc_bench.txt 2>&1 <<'PY'
import importlib, json, os, sys, torch
sys.path.insert(0, os.getcwd())
E = importlib.import_module("rag-project-icd10_b200.engine.encoder")
peaks = {"bf16_tflops_sustained": 1399.4}
dev = torch.device("cuda", 0)
for S, B in ((64, 4096), (32, 8192), (128, 2048), (24, 4096)):
    print(S, B, json.dumps(E.bench_encoder(dev, peaks, batch=B, seq=S, steps=5, warmup=2)))
PY
cat $OUT/enc_bench.txt
ICD_ATTN_CUDA_CORE=1 python - > $OUT/enc_bench_oldattn.txt 2>&1 <<'PY'
import importlib, json, os, sys, torch
sys.path.insert(0, os.getcwd())
E = importlib.import_module("rag-project-icd10_b200.engine.encoder")
print(json.dumps(E.bench_encoder(torch.device("cuda", 0), {"bf16_tflops_sustained": 1399.4}, steps=5, warmup=2)))
PY
cat $OUT/enc_bench_oldattn.txt
ENC_REPS=2 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'gemm_tc|attention|layernorm|embed|pool' -s 86 -c 8 --csv --log-file $OUT/enc_layer_times.csv python profiles/encoder_once.py > /dev/null 2>&1
cut -d, -f5,15 $OUT/enc_layer_times.csv | tail -n 9
