"""Worker of the sharded-build test: run under torch.distributed.run, one rank per GPU.  Builds the database with the
records split over the ranks (tools/build_database.py) and checks it against a single-GPU build."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    work = os.environ["ICD_TEST_WORKDIR"]
    csv_path = os.path.join(work, "subset.csv")
    os.environ["EMBEDDING_MODEL_NAME"] = os.path.join(work, "model")
    os.environ["MILVUS_DB_PATH"] = os.path.join(work, "db", "sharded.db")
    os.environ["MILVUS_COLLECTION_NAME"] = "icd10"
    os.chdir(work)
    B = importlib.import_module("rag-project-icd10_b200.tools.build_database")
    M = importlib.import_module("rag-project-icd10_b200.services.milvus_service")
    builder = B.DatabaseBuilder(rank=rank, world=world, dist=dist)
    ok = builder.build_full_database(csv_path, rebuild=True)
    es, ms = builder.embedding_service, builder.milvus_service
    n = ms.get_collection_stats()["num_entities"]
    col = ms.client.cols["icd10"]
    lo, hi = col.row_lo, col.row_hi
    probes = ["急性胃肠炎", "霍乱", "伤寒", "结核性脑膜炎", "急性胃肠炎 发热"]
    sharded = [ms.search(es.encode_query(p), top_k=10) for p in probes]               # collective: every rank
    # the batched call against single calls on the SAME vectors (a text encoded alone takes the few-token kernels, in a
    # batch the tile kernels: equal to bf16 rounding, which may swap near-ties of these random-weight embeddings)
    qv = es.encode_queries(probes)
    batch = [list(c) for c in ms.search_batch(qv, top_k=10)]
    one_by_one = [ms.search(qv[i], top_k=10) for i in range(len(probes))]
    res = {"ok": bool(ok), "n": n, "rows": [lo, hi], "resident": len(col.index)}
    good = ok and n == int(os.environ["ICD_TEST_ROWS"]) and len(col.index) == hi - lo
    good = good and all([h["code"] for h in a] == [h["code"] for h in b] for a, b in zip(one_by_one, batch))
    good = good and all(len({h["code"] for h in a} & {h["code"] for h in b}) >= 8 for a, b in zip(sharded, batch))
    if not good:
        print("rank", rank, "batched vs single searches differ", flush=True)
    # a fresh sharded service maps its slice back from the column files and answers identically
    ms2 = M.MilvusService(embedding_service=es)
    shard = importlib.import_module("rag-project-icd10_b200.engine.shard")
    c2 = ms2.client.cols["icd10"]
    grp = shard.ShardGroup(c2.index, row_offset=c2.row_lo, rank=rank, world=world)
    ms2.client.attach_group("icd10", grp)
    again = [ms2.search(es.encode_query(p), top_k=10) for p in probes]
    good = good and (c2.row_lo, c2.row_hi) == (lo, hi) and again == sharded
    dist.barrier()
    if rank == 0:
        # single-GPU build of the same CSV, unsharded client
        M.MilvusService.client_kwargs = {}
        os.environ["MILVUS_DB_PATH"] = os.path.join(work, "db", "single.db")
        single = B.DatabaseBuilder()
        assert single.build_full_database(csv_path, rebuild=True)
        want = [single.milvus_service.search(single.embedding_service.encode_query(p), top_k=10) for p in probes]
        for a, b in zip(sharded, want):
            sa = {h["code"]: h for h in a}
            common = [h for h in b if h["code"] in sa]
            good = good and len(common) >= 9 and all(abs(h["original_score"] - sa[h["code"]]["original_score"]) < 1e-4 and
                                                      h["metadata"] == sa[h["code"]]["metadata"] for h in common)
        d1 = os.path.join(work, "db", "sharded.db.icdb", "icd10")
        d2 = os.path.join(work, "db", "single.db.icdb", "icd10")
        for fn in sorted(os.listdir(d2)):
            a, b = open(os.path.join(d1, fn), "rb").read(), open(os.path.join(d2, fn), "rb").read()
            if fn == "vectors.f32":
                va, vb = np.frombuffer(a, "<f4").reshape(n, -1), np.frombuffer(b, "<f4").reshape(n, -1)
                cos = (va * vb).sum(1)
                good = good and bool(cos.min() > 0.99999)
                res["min_cos_sharded_vs_single"] = float(cos.min())
            elif fn == "vectors.bf16":
                good = good and len(a) == len(b)
            else:
                if a != b:
                    print("column file differs:", fn, flush=True)
                good = good and a == b
        single.milvus_service.disconnect()
    flag = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(json.dumps({"rank": rank, **res}), flush=True)
    grp.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("BUILD_OK" if int(flag) == 1 else "BUILD_FAIL", flush=True)
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
