"""Batch-1 encode latency: few-token weight-streaming path (default) vs the tile kernels (enc_skinny = 0), 12 layers."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
N = importlib.import_module("rag-project-icd10_b200._native")
eng = bench.synthetic_engine(num_layers=12, device=0, max_tokens=8192)
rng = np.random.default_rng(1)
for S in (8, 12, 16, 17, 24, 32, 48, 64, 65):
    ids = rng.integers(1000, 20000, size=(1, S)).astype(np.int32); lens = np.array([S], np.int32)
    d_ids = torch.from_numpy(ids).cuda(); d_lens = torch.from_numpy(lens).cuda()
    row = {"S": S}
    for mode in (1, 0, 1, 0):
        N.tune(enc_skinny=mode)
        for _ in range(20):
            eng.forward_ids(ids, lens)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); eng.forward_ids(ids, lens); ts.append(time.perf_counter() - t0)
        ts.sort()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            eng.forward_ids(d_ids, d_lens)
        e1.record(); torch.cuda.synchronize()
        row[f"skinny{mode}_host_p50_ms"] = round(ts[100] * 1e3, 4)
        row[f"skinny{mode}_host_p99_ms"] = round(ts[197] * 1e3, 4)
        row[f"skinny{mode}_device_ms"] = round(e0.elapsed_time(e1) / 100, 4)
    print(json.dumps(row), flush=True)
N.tune(enc_skinny=1)
