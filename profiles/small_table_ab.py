"""Small-table regime of the tensor scan (BASELINE configs[1]: 40 474 rows): pre-pass on (default) vs off
(icd_tune scan_small_pre = 0, the round-2 behaviour), same handle, alternating; results must be identical."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
N = importlib.import_module("rag-project-icd10_b200._native")
VectorIndex = importlib.import_module("rag-project-icd10_b200.engine.index").VectorIndex
rng = np.random.default_rng(5)
nq = 10000
q = rng.standard_normal((nq, 768)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
sizes = [int(a) for a in sys.argv[1:]] or [40474, 2048, 300000]
for n in sizes:
    corpus = rng.standard_normal((n, 768)).astype(np.float32); corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    levels = rng.integers(1, 4, size=n).astype(np.uint8)
    for keep in (True, False):
        idx = VectorIndex(768, device=0, keep_f32=keep)
        idx.append(corpus, levels)
        idx.set_timing(True)
        for B in (8, 64, 256, 1024, 8192, 10000):
            row = {"rows": n, "keep_f32": keep, "B": B}
            res = {}
            for pre in (0, 1, 0, 1):
                N.tune(scan_small_pre=pre)
                idx.search(q[:B], 10)
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter(); r = idx.search(q[:B], 10); ts.append(time.perf_counter() - t0)
                res[pre] = r
                tm = idx.last_timing()
                row[f"pre{pre}_ms"] = round(min(ts) * 1e3, 3)
                row[f"pre{pre}_scan_us"] = round(tm["scan_us"], 1)
                row[f"pre{pre}_merge_us"] = round(tm["merge_us"], 1)
                row[f"pre{pre}_finalise_us"] = round(tm["finalise_us"], 1)
                row[f"pre{pre}_launches"] = tm["launches"]
            row["identical"] = bool(all(np.array_equal(a, b) for a, b in zip(res[0], res[1])))
            print(json.dumps(row), flush=True)
        N.tune(scan_small_pre=1)
        idx.close()
