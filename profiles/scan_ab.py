"""Same-box A/B of scan knobs on one resident table: builds the synthetic corpus once, then times each knob setting
(CUDA events, STEPS steps after 2 warm-ups), twice in alternation.  ROWS / BATCH / VARIANTS from the environment:
  ROWS=100000000 BATCH=1024 VARIANTS="scan_pair=-1;scan_pair=0;scan_pair=0,scan_qsplit=0" python profiles/scan_ab.py"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
N = importlib.import_module("rag-project-icd10_b200._native")
if os.environ.get("PROF_LIB"):      # a profiling build of the library (make PROFILING=1), e.g. for layout experiments
    N.LIB_PATH = os.path.abspath(os.environ["PROF_LIB"])
VectorIndex = importlib.import_module("rag-project-icd10_b200.engine.index").VectorIndex
rows, B, k = int(os.environ.get("ROWS", 100_000_000)), int(os.environ.get("BATCH", 1024)), 10
steps = int(os.environ.get("STEPS", 5))
warm = int(os.environ.get("WARM", 2))
variants = [v for v in os.environ.get("VARIANTS", "scan_pair=-1;scan_pair=0").split(";") if v]
DEFAULTS = dict(scan_sample=-1, scan_drift=4, scan_tmax=16, scan_kbs=3, scan_kbs_pair=6, scan_qsplit=-1, scan_pair=-1, scan_qtmem=0,
                scan_generic=0, scan_pre_slots=1)
if os.environ.get("PROF_LIB") and os.environ.get("PROF_TILED"):   # profiling builds only (make PROFILING=1)
    DEFAULTS["scan_tiled"] = 0
dev = torch.device("cuda", 0)
table, levels = bench.make_corpus(torch, rows, dev, 1234)
q = bench.make_queries(torch, B, dev)
idx = VectorIndex(768, device=0)
idx.adopt(table, levels)
out = (torch.empty((B, k), dtype=torch.float32, device=dev), torch.empty((B, k), dtype=torch.float32, device=dev),
       torch.empty((B, k), dtype=torch.int64, device=dev))
st = torch.cuda.current_stream(dev)
ref_ids = None
for rep in range(2):
    for v in variants:
        N.tune(**DEFAULTS)
        N.tune(**{kv.split("=")[0]: int(kv.split("=")[1]) for kv in v.split(",")})
        for _ in range(warm):
            idx.search(q, k, out=out, stream=st.cuda_stream, sync=False)
        torch.cuda.synchronize()
        idx.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            idx.search(q, k, out=out, stream=st.cuda_stream, sync=False)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        tm = idx.mean_timing()
        ids = out[2].clone()
        same = True if ref_ids is None else bool(torch.equal(ids, ref_ids))
        ref_ids = ids if ref_ids is None else ref_ids
        tf = 2.0 * B * rows * 768 / (tm["scan_us"] * 1e-6) / 1e12
        print(json.dumps({"rows": rows, "batch": B, "tune": v, "rep": rep, "ms_per_step": round(ms, 3), "qps": round(B / ms * 1e3),
                          "scan_us": round(tm["scan_us"]), "tflops": round(tf, 1), "ids_same_as_first": same}), flush=True)
