"""Times the encoder alone (B=4096, S=64, 12 layers, synthetic): ENC_REPS timed batches after 3 warm-ups."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as E
import importlib as _il
_il.import_module("rag-project-icd10_b200._native").tune(enc_pdl=int(os.environ.get("ENC_PDL", "0")))
B, S = int(os.environ.get("ENC_B", 4096)), int(os.environ.get("ENC_S", 64))
reps = int(os.environ.get("ENC_REPS", 20))
eng = E.synthetic_engine(device=0, max_tokens=B * S)
ids = torch.randint(1000, 21128, (B, S), device="cuda", dtype=torch.int32)
lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
out = torch.empty((B, 768), dtype=torch.float32, device="cuda")
for _ in range(3):
    eng.forward_ids(ids, lens, out=out)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    eng.forward_ids(ids, lens, out=out)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
# same-box cuBLAS reference (box-to-box clocks under the power cap differ by a few %): sustained bf16 GEMM, ~1 s
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(20): a @ b
torch.cuda.synchronize(); t0.record()
for _ in range(600): a @ b
t1.record(); torch.cuda.synchronize()
cublas_tf = 600 * 2 * 8192 ** 3 / (t0.elapsed_time(t1) * 1e-3) / 1e12
flops = B * S * 12 * (2 * (4 * 768 * 768 + 2 * 768 * 3072) + 4 * S * 768)
print({"fused_ln": os.environ.get("ICD_ENC_FUSED_LN", "1"), "pair": os.environ.get("ICD_GEMM_PAIR", "1"), "B": B, "S": S,
       "ms_per_batch": round(ms, 3), "sentences_per_s": round(B / ms * 1e3), "tflops": round(flops / ms / 1e9, 1),
       "mean_norm": float(out.norm(dim=1).mean()), "cublas_sustained_tflops_this_box": round(cublas_tf, 1),
       "frac_of_this_box_cublas": round(flops / ms / 1e9 / cublas_tf, 3)})
