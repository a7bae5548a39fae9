"""Oracle: Milvus FLAT/IP search + level re-rank (test infrastructure only).

Third-party engine restated: pymilvus==2.5.10 / milvus-lite brute-force FLAT index with
metric IP (/root/reference/requirements.txt:35; configured at
/root/reference/services/milvus_service.py:33-34,189-194): fp32 inner product of the query
with every stored vector, top-`limit` by descending score.  The re-rank, hit layout and
re-sort follow /root/reference/services/milvus_service.py:280-314,550-558.
Parity unpinned for the engine arithmetic (no fixtures in the reference); ties are ordered
(score desc, row id asc), which is what a stable sort over insertion order gives.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .text import level_weight

OUTPUT_FIELDS = ["code", "preferred_zh", "has_complication", "main_code", "secondary_code",
                 "level", "parent_code", "category_path", "semantic_text"]


def exact_topk(corpus: np.ndarray, queries: np.ndarray, k: int, row_block: int = 1 << 18):
    """corpus [N,D] (any float dtype, up-cast to fp32), queries [B,D] fp32.

    Returns (scores [B,k'] fp32, ids [B,k'] int64), k' = min(k, N), ordered by
    (score desc, id asc).  Blocked over rows so a 10 M-row corpus stays in cache-sized
    pieces; selection is exact (lexsort on the candidate pool).
    """
    corpus = np.asarray(corpus)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    n = corpus.shape[0]
    b = q.shape[0]
    kk = min(k, n)
    if kk == 0:
        return np.zeros((b, 0), np.float32), np.zeros((b, 0), np.int64)
    pool_s = np.full((b, 0), -np.inf, np.float32)
    pool_i = np.zeros((b, 0), np.int64)
    for lo in range(0, n, row_block):
        hi = min(n, lo + row_block)
        blk = np.asarray(corpus[lo:hi], dtype=np.float32)
        s = q @ blk.T                                    # [b, hi-lo] fp32
        ids = np.broadcast_to(np.arange(lo, hi, dtype=np.int64), s.shape)
        if s.shape[1] > 4 * kk:
            # keep everything >= the kk-th value of this block (ties included)
            part = np.partition(s, s.shape[1] - kk, axis=1)[:, s.shape[1] - kk]
            keep_s, keep_i = [], []
            width = 0
            for r in range(b):
                m = s[r] >= part[r]
                keep_s.append(s[r][m]); keep_i.append(ids[r][m]); width = max(width, int(m.sum()))
            bs = np.full((b, width), -np.inf, np.float32); bi = np.full((b, width), np.iinfo(np.int64).max, np.int64)
            for r in range(b):
                bs[r, :keep_s[r].size] = keep_s[r]; bi[r, :keep_i[r].size] = keep_i[r]
            s, ids = bs, bi
        pool_s = np.concatenate([pool_s, s], axis=1)
        pool_i = np.concatenate([pool_i, ids], axis=1)
        if pool_s.shape[1] > 8 * kk:
            pool_s, pool_i = _cut(pool_s, pool_i, kk)
    return _cut(pool_s, pool_i, kk)


def fast_topk(corpus: np.ndarray, queries: np.ndarray, k: int, row_block: int = 1 << 17):
    """The timed CPU leg of bench.py (cpu_baseline / --impl reference): the same exact fp32
    inner-product search as exact_topk -- BLAS GEMM per row block, then a fully vectorised
    argpartition instead of exact_topk's per-query tie bookkeeping -- so the CPU arm is not
    slowed down by Python loops.  Agrees with exact_topk whenever the k-th and (k+1)-th scores of a
    block differ (tests/test_oracle_golden.py)."""
    corpus = np.asarray(corpus)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    n, b = corpus.shape[0], q.shape[0]
    kk = min(k, n)
    pool_s, pool_i = [], []
    for lo in range(0, n, row_block):
        hi = min(n, lo + row_block)
        s = q @ np.asarray(corpus[lo:hi], dtype=np.float32).T
        if s.shape[1] > kk:
            part = np.argpartition(s, s.shape[1] - kk, axis=1)[:, s.shape[1] - kk:]
            pool_s.append(np.take_along_axis(s, part, axis=1))
            pool_i.append(part.astype(np.int64) + lo)
        else:
            pool_s.append(s)
            pool_i.append(np.broadcast_to(np.arange(lo, hi, dtype=np.int64), s.shape))
    ps, pi = np.concatenate(pool_s, axis=1), np.concatenate(pool_i, axis=1)
    order = np.lexsort((pi, -ps.astype(np.float64)), axis=1)[:, :kk]
    return np.take_along_axis(ps, order, axis=1), np.take_along_axis(pi, order, axis=1)


def _cut(s: np.ndarray, i: np.ndarray, k: int):
    out_s = np.empty((s.shape[0], k), np.float32)
    out_i = np.empty((s.shape[0], k), np.int64)
    for r in range(s.shape[0]):
        order = np.lexsort((i[r], -s[r].astype(np.float64)))[:k]
        out_s[r] = s[r][order]
        out_i[r] = i[r][order]
    return out_s, out_i


def rerank(raw: Sequence[float], levels: Sequence[int]):
    """milvus_service.py:290-314: score = float(raw)*weight(level); stable sort desc by score.

    Returns (order, weighted) where order indexes the input hits."""
    weighted = [float(float(r) * level_weight(int(l))) for r, l in zip(raw, levels)]
    order = sorted(range(len(weighted)), key=lambda j: weighted[j], reverse=True)
    return order, weighted


def search_hits(corpus: np.ndarray, records: List[dict], query_vector: np.ndarray, top_k: int = 10) -> List[dict]:
    """What MilvusService.search returns (milvus_service.py:271-320) for one query."""
    scores, ids = exact_topk(corpus, np.asarray(query_vector, np.float32)[None, :], top_k)
    hits = []
    for s, j in zip(scores[0], ids[0]):
        rec = records[int(j)]
        base = float(s)
        level = rec.get("level", 1)
        hits.append({
            "code": rec.get("code"),
            "title": rec.get("preferred_zh"),
            "score": float(base * level_weight(level)),
            "original_score": base,
            "metadata": {
                "has_complication": rec.get("has_complication", False),
                "main_code": rec.get("main_code", ""),
                "secondary_code": rec.get("secondary_code", ""),
                "level": level,
                "parent_code": rec.get("parent_code", ""),
                "category_path": rec.get("category_path", ""),
                "semantic_text": rec.get("semantic_text", ""),
            },
        })
    hits.sort(key=lambda h: h["score"], reverse=True)
    return hits


def merge_shards(shard_scores: Sequence[np.ndarray], shard_ids: Sequence[np.ndarray], k: int):
    """Merge per-shard [B,k] candidates (global ids) by (score desc, id asc)."""
    s = np.concatenate(list(shard_scores), axis=1)
    i = np.concatenate(list(shard_ids), axis=1)
    return _cut(s.astype(np.float32), i.astype(np.int64), min(k, s.shape[1]))


# ---------------------------------------------------------------- synthetic data
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (nearest even) and return as fp32 holding bf16-representable values."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    return (bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


def from_bf16_bits(u16: np.ndarray) -> np.ndarray:
    return (np.asarray(u16, np.uint16).astype(np.uint32) << 16).view(np.float32)
