"""Where does MilvusService.search_batch(10 000 queries) spend its time?  (configs[1] through the service API)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
N = importlib.import_module("rag-project-icd10_b200._native")
VectorIndex = importlib.import_module("rag-project-icd10_b200.engine.index").VectorIndex
rng = np.random.default_rng(5)
n, nq = 40474, 10000
corpus = rng.standard_normal((n, 768)).astype(np.float32); corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
levels = rng.integers(1, 4, size=n).astype(np.uint8)
q = rng.standard_normal((nq, 768)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
for keep in (True, False):
    idx = VectorIndex(768, device=0, keep_f32=keep)
    idx.append(corpus, levels)
    idx.set_timing(True)
    for B in (64, 1024, 8192, 10000):
        idx.search(q[:B], 10)
        t0 = time.perf_counter(); idx.search(q[:B], 10); dt = time.perf_counter() - t0
        print("keep_f32", keep, "B", B, "ms", round(dt * 1e3, 2), idx.last_timing(), flush=True)
    idx.close()
