"""Worker for the multi-GPU shard tests: run under torch.distributed.run, one rank per GPU.
Checks that the row-sharded search (both exchanges) returns exactly what a single table does."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import search as osearch
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = importlib.import_module("rag-project-icd10_b200._native")
    VectorIndex = importlib.import_module("rag-project-icd10_b200.engine.index").VectorIndex
    shard = importlib.import_module("rag-project-icd10_b200.engine.shard")

    n, dim, B, k = 60_001, 768, 150, 10
    rng = np.random.default_rng(77)          # same data on every rank
    corpus = rng.standard_normal((n, dim)).astype(np.float32)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    corpus = osearch.bf16_round(corpus)
    corpus[40_000] = corpus[5]                # a cross-shard exact tie
    levels = rng.integers(1, 4, size=n).astype(np.uint8)
    q = osearch.bf16_round(corpus[rng.integers(0, n, size=B)] + 0.02 * rng.standard_normal((B, dim)).astype(np.float32))
    q[0] = corpus[5]
    lo, hi = shard.shard_bounds(n, rank, world)
    idx = VectorIndex(dim, device=local)
    idx.append(corpus[lo:hi], levels[lo:hi])
    grp = shard.ShardGroup(idx, row_offset=lo, rank=rank, world=world)

    whole = VectorIndex(dim, device=local)
    whole.append(corpus, levels)
    ok = True
    for mode in (N.WEIGHT_RERANK, N.WEIGHT_NONE, N.WEIGHT_PRE):
        ws, wr, wi = whole.search(q, k, weight_mode=mode, path=N.PATH_TENSOR)
        for exchange in (0, 1):
            for path in (N.PATH_STREAM, N.PATH_TENSOR):
                qq = q[:6] if path == N.PATH_STREAM else q
                s, r, i = grp.search(qq, k, weight_mode=mode, path=path, exchange=exchange)
                m = len(qq)
                good = np.array_equal(i, wi[:m]) and np.array_equal(r, wr[:m]) and np.array_equal(s, ws[:m])
                if not good:
                    bad = np.nonzero((i != wi[:m]).any(1))[0][:3]
                    print(f"rank {rank} MISMATCH mode={mode} exchange={exchange} path={path} queries={bad} "
                          f"{i[bad[:1]]} vs {wi[bad[:1]]}", flush=True)
                ok = ok and good
    assert wi[0, 0] == 5 and wi[0, 1] == 40_000, wi[0]
    # repeated calls exercise the slab parity/epoch protocol; device buffers + async stream
    qd = torch.from_numpy(q).cuda().to(torch.bfloat16)
    for it in range(20):
        s, r, i = grp.search(qd, k, exchange=1, sync=False, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ws, wr, wi = whole.search(q, k, path=N.PATH_TENSOR)
    ok = ok and np.array_equal(i.cpu().numpy(), wi)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    grp.close(); idx.close(); whole.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARD_OK" if int(flag) == 1 else "SHARD_FAIL", flush=True)
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
