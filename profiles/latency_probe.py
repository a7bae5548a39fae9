"""Batch-1 encode latency: wall vs GPU-side (CUDA events) for a 12-layer synthetic encoder, S = 12 / 24 / 48 tokens."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
N = importlib.import_module("rag-project-icd10_b200._native")
N.tune(enc_pdl=int(os.environ.get("ENC_PDL", "0")))
print("enc_pdl", os.environ.get("ENC_PDL", "0"))
eng = bench.synthetic_engine(device=0, max_tokens=8192)
st = torch.cuda.current_stream()
for S in (12, 24, 48, 128):
    ids = np.random.randint(1000, 20000, (1, S)).astype(np.int32); lens = np.array([S], np.int32)
    d_ids, d_lens = torch.from_numpy(ids).cuda(), torch.from_numpy(lens).cuda()
    d_out = torch.empty((1, 768), dtype=torch.float32, device="cuda")
    for _ in range(20):
        eng.forward_ids(ids, lens)
    t0 = time.perf_counter()
    for _ in range(500):
        eng.forward_ids(ids, lens)
    wall = (time.perf_counter() - t0) / 500 * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(500):
        eng.forward_ids(d_ids, d_lens, out=d_out, stream=st.cuda_stream, sync=False)
    e1.record(st); torch.cuda.synchronize()
    gpu = e0.elapsed_time(e1) / 500
    print(f"S={S}: host-buffer call {wall:.3f} ms; back-to-back device-buffer launches {gpu:.3f} ms per forward (63 launches)", flush=True)
