"""One configs[1]-shaped search (40 474 rows, B from $B, default 8192) after two warm-ups: the target of an ncu capture."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
VectorIndex = importlib.import_module("rag-project-icd10_b200.engine.index").VectorIndex
B = int(os.environ.get("B", "8192")); n = int(os.environ.get("ROWS", "40474"))
rng = np.random.default_rng(5)
corpus = rng.standard_normal((n, 768)).astype(np.float32); corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
levels = rng.integers(1, 4, size=n).astype(np.uint8)
q = rng.standard_normal((B, 768)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
idx = VectorIndex(768, device=0, keep_f32=True)
idx.append(corpus, levels)
for _ in range(3):
    idx.search(q, 10)
idx.close()
