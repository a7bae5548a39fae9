"""One resident corpus, several batch sizes, ONE profiled search each (cudaProfilerStart/Stop around it) -- for
`ncu --profile-from-start off --set full -k regex:scan_tc`:   ROWS=12500000 BATCHES=1024,256,128 python profiles/ncu_scan.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
N = importlib.import_module("rag-project-icd10_b200._native")
VectorIndex = importlib.import_module("rag-project-icd10_b200.engine.index").VectorIndex
rows = int(os.environ.get("ROWS", 10_000_000))
batches = [int(b) for b in os.environ.get("BATCHES", "1024").split(",")]
dev = torch.device("cuda", 0)
table, levels = bench.make_corpus(torch, rows, dev, 1234)
q_all, _, _ = bench.make_planted_queries(torch, None, table, 0, rows, rows, max(batches), dev, 0, 1)
idx = VectorIndex(768, device=0)
idx.adopt(table, levels)
for B in batches:
    q = q_all[:B].contiguous()
    out = (torch.empty((B, 10), dtype=torch.float32, device=dev), torch.empty((B, 10), dtype=torch.float32, device=dev),
           torch.empty((B, 10), dtype=torch.int64, device=dev))
    idx.search(q, 10, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    idx.search(q, 10, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", rows, B, flush=True)
