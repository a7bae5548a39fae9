"""One encoder batch (B=4096, S=64, 12 layers, synthetic) for ncu launch lists."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as E
B, S = int(os.environ.get("ENC_B", 4096)), int(os.environ.get("ENC_S", 64))
eng = E.synthetic_engine(device=0, max_tokens=B * S)
ids = torch.randint(1000, 21128, (B, S), device="cuda", dtype=torch.int32)
lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
out = torch.empty((B, 768), dtype=torch.float32, device="cuda")
for _ in range(int(os.environ.get("ENC_REPS", 2))):
    eng.forward_ids(ids, lens, out=out)
torch.cuda.synchronize()
print("ok", float(out.norm(dim=1).mean()))
