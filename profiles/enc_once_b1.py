"""Three batch-1 forwards of S tokens (env S, default 12) through the 12-layer synthetic encoder: ncu launch-list target."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
N = importlib.import_module("rag-project-icd10_b200._native")
if os.environ.get("SKINNY"):
    N.tune(enc_skinny=int(os.environ["SKINNY"]))
eng = bench.synthetic_engine(num_layers=12, device=0, max_tokens=8192)
S = int(os.environ.get("S", "12"))
ids = np.random.default_rng(1).integers(1000, 20000, size=(1, S)).astype(np.int32)
for _ in range(3):
    eng.forward_ids(ids, np.array([S], np.int32))
