"""Host-side logic of the row-sharded search, on CPU: shard bounds, the merge rule the GPU
exchange implements (oracle level), and the torch.distributed bootstrap under gloo, world 2."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from oracle import search as osearch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    shard = importlib.import_module("rag-project-icd10_b200.engine.shard")
    for total in (0, 1, 7, 40474, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_of_shard_topk_equals_global_topk():
    rng = np.random.default_rng(9)
    c = rng.standard_normal((9000, 64)).astype(np.float32)
    c[7000] = c[12]
    q = np.concatenate([c[12:13], rng.standard_normal((5, 64)).astype(np.float32)])
    ref_s, ref_i = osearch.exact_topk(c, q, 10)
    shard = importlib.import_module("rag-project-icd10_b200.engine.shard")
    for world in (2, 4, 8):
        ss, ii = [], []
        for r in range(world):
            lo, hi = shard.shard_bounds(len(c), r, world)
            s, i = osearch.exact_topk(c[lo:hi], q, 10)
            ss.append(s); ii.append(i + lo)
        m_s, m_i = osearch.merge_shards(ss, ii, 10)
        assert np.array_equal(m_i, ref_i) and np.allclose(m_s, ref_s, atol=1e-6)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("rag-project-icd10_b200.engine.shard")
    ident, table = shard.bootstrap_exchange(dist, rank, world, (bytes([rank + 1]) * 128) if rank == 0 else None,
                                            bytes([10 + rank]) * 64, share_ident=True)
    # each rank searches its rows with the oracle; merged result must equal the global one
    rng = np.random.default_rng(4)
    c = rng.standard_normal((2001, 32)).astype(np.float32)
    qs = rng.standard_normal((4, 32)).astype(np.float32)
    lo, hi = shard.shard_bounds(len(c), rank, world)
    s, i = osearch.exact_topk(c[lo:hi], qs, 5)
    gathered = [None] * world
    dist.all_gather_object(gathered, (s, i + lo))
    m_s, m_i = osearch.merge_shards([g[0] for g in gathered], [g[1] for g in gathered], 5)
    ref_s, ref_i = osearch.exact_topk(c, qs, 5)
    q.put((rank, ident == bytes([1]) * 128, table == bytes([10]) * 64 + bytes([11]) * 64, bool(np.array_equal(m_i, ref_i))))
    dist.destroy_process_group()


def test_bootstrap_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_partition_by_tokens_balances_and_covers():
    shard = importlib.import_module("rag-project-icd10_b200.engine.shard")
    rng = np.random.default_rng(3)
    counts = rng.integers(3, 129, size=1001).tolist()
    for world in (1, 2, 3, 8):
        parts = shard.partition_by_tokens(counts, world)
        assert sorted(j for p in parts for j in p) == list(range(len(counts)))
        loads = [sum(counts[j] for j in p) for p in parts]
        assert max(loads) - min(loads) <= 128          # within one longest sentence
    assert shard.partition_by_tokens([], 4) == [[], [], [], []]
    assert shard.partition_by_tokens([5], 2) == [[0], []]


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("rag-project-icd10_b200.engine.shard")
    texts = ["t%03d" % i + "x" * (i % 17) for i in range(53)]
    counts = [len(t) + 2 for t in texts]

    def fake_encode(batch):           # a deterministic stand-in for an encoder replica
        return np.stack([np.full(8, float(int(t[1:4])), np.float32) for t in batch]) if batch else np.zeros((0, 8), np.float32)
    full = shard.encode_data_parallel(fake_encode, texts, counts, rank, world, dist=dist, dim=8)
    local = shard.encode_data_parallel(fake_encode, texts, counts, rank, world, dist=None, dim=8)
    mine = shard.partition_by_tokens(counts, world)[rank]
    ok_full = bool(np.array_equal(full[:, 0], np.arange(53, dtype=np.float32)))
    ok_local = bool(np.array_equal(local[mine, 0], np.asarray(mine, np.float32))) and float(np.abs(local).sum()) == float(sum(mine)) * 8
    q.put((rank, ok_full, ok_local))
    dist.destroy_process_group()


def test_encode_data_parallel_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
