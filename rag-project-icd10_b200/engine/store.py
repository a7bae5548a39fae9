"""IcdStoreClient -- the object MilvusService holds where the reference holds a
``pymilvus.MilvusClient`` (/root/reference/services/milvus_service.py:81,100-110).

It offers the slice of the MilvusClient surface the reference touches (has_collection,
create_schema / add_field, prepare_index_params / add_index, create_collection,
get_load_state, load_collection, release_collection, drop_collection, get_collection_stats,
insert, search, close) over:
  * a device-resident vector table searched by libicdrag.so (engine/index.py): bf16 scan copy
    plus an fp32 master, so single-query search is exact fp32 inner product like Milvus FLAT/IP;
  * host-side columns for the scalar fields of the schema (milvus_service.py:163-206);
  * an append-only on-disk form next to MILVUS_DB_PATH: <db_path>.icdb/<collection>.{json,vec,meta}
    (vectors as raw little-endian float32 rows, metadata as JSON lines) with the reference's
    append / drop semantics (auto-id primary key, re-running the build appends duplicates).
There is no CPU search path: loading a collection needs a GPU.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

from .. import _native as N
from .index import VectorIndex

SCALAR_FIELDS = ["code", "preferred_zh", "has_complication", "main_code", "secondary_code", "level",
                 "parent_code", "category_path", "semantic_text"]


class Hit(dict):
    """pymilvus Hit look-alike: ``hit.get(field)`` falls through to the entity fields
    (the reference relies on it at milvus_service.py:293,298-309)."""

    def get(self, key, default=None):
        if key in self:
            return dict.get(self, key)
        return dict.get(self, "entity", {}).get(key, default)


class _Schema:
    def __init__(self, **kw):
        self.options = kw
        self.fields: List[Dict[str, Any]] = []

    def add_field(self, field_name: str, datatype=None, **kw):
        self.fields.append({"field_name": field_name, "datatype": datatype, **kw})
        return self

    @property
    def dim(self) -> Optional[int]:
        for f in self.fields:
            if "dim" in f:
                return int(f["dim"])
        return None


class _IndexParams:
    def __init__(self):
        self.indexes: List[Dict[str, Any]] = []

    def add_index(self, field_name: str, index_type: str = "FLAT", metric_type: str = "IP", **kw):
        self.indexes.append({"field_name": field_name, "index_type": index_type, "metric_type": metric_type, **kw})


class DataType:
    """Names the reference imports from pymilvus (milvus_service.py:6,174-186)."""
    INT64, FLOAT_VECTOR, VARCHAR, BOOL, INT32 = "INT64", "FLOAT_VECTOR", "VARCHAR", "BOOL", "INT32"


class _Collection:
    def __init__(self, name: str, dim: int, metric: str = "IP", index_type: str = "FLAT"):
        self.name, self.dim, self.metric, self.index_type = name, dim, metric, index_type
        self.rows: List[Dict[str, Any]] = []       # scalar fields, insertion order == row id
        self.levels = np.zeros((0,), np.uint8)
        self.index: Optional[VectorIndex] = None   # device table when loaded
        self.loaded = False


class IcdStoreClient:
    def __init__(self, uri: str = "./db/milvus_icd10.db", device: int = 0, **_remote_kwargs):
        N.require_gpu()
        self.uri, self.device = uri, device
        self.root = uri + ".icdb" if not uri.endswith(".icdb") else uri
        self.persist = not uri.startswith(("http://", "https://"))
        self.cols: Dict[str, _Collection] = {}
        if self.persist:
            os.makedirs(self.root, exist_ok=True)
            for fn in sorted(os.listdir(self.root)):
                if fn.endswith(".json"):
                    self._open(fn[:-5])

    # ---------------------------------------------------------------- persistence
    def _paths(self, name: str):
        base = os.path.join(self.root, name)
        return base + ".json", base + ".vec", base + ".meta"

    def _open(self, name: str) -> None:
        head, vec, meta = self._paths(name)
        with open(head, encoding="utf-8") as fh:
            h = json.load(fh)
        col = _Collection(name, int(h["dim"]), h.get("metric", "IP"), h.get("index_type", "FLAT"))
        if os.path.exists(meta):
            with open(meta, encoding="utf-8") as fh:
                col.rows = [json.loads(line) for line in fh if line.strip()]
        col.levels = np.asarray([_level_byte(r.get("level", 1)) for r in col.rows], np.uint8)
        self.cols[name] = col

    def _vectors_from_disk(self, col: _Collection) -> np.ndarray:
        _, vec, _ = self._paths(col.name)
        if not self.persist or not os.path.exists(vec):
            return np.zeros((0, col.dim), np.float32)
        data = np.fromfile(vec, dtype="<f4")
        n = len(col.rows)
        if data.size < n * col.dim:
            raise N.NativeError(f"collection {col.name}: vector file is shorter than its metadata")
        return data[: n * col.dim].reshape(n, col.dim)

    # ---------------------------------------------------------------- MilvusClient surface
    def has_collection(self, collection_name: str) -> bool:
        return collection_name in self.cols

    def create_schema(self, **kw) -> _Schema:
        return _Schema(**kw)

    def prepare_index_params(self) -> _IndexParams:
        return _IndexParams()

    def create_collection(self, collection_name: str, schema: _Schema = None, index_params: _IndexParams = None,
                          dimension: Optional[int] = None, **_kw) -> None:
        dim = dimension or (schema.dim if schema is not None else None)
        if not dim:
            raise ValueError("create_collection needs a FLOAT_VECTOR field with dim")
        metric, itype = "IP", "FLAT"
        if index_params is not None and index_params.indexes:
            metric = index_params.indexes[0].get("metric_type", "IP")
            itype = index_params.indexes[0].get("index_type", "FLAT")
        if metric != "IP":
            raise N.NativeError(f"only the IP metric is implemented (got {metric})")
        col = _Collection(collection_name, int(dim), metric, itype)
        self.cols[collection_name] = col
        if self.persist:
            head, vec, meta = self._paths(collection_name)
            with open(head, "w", encoding="utf-8") as fh:
                json.dump({"dim": col.dim, "metric": metric, "index_type": itype,
                           "fields": [f["field_name"] for f in (schema.fields if schema else [])]}, fh)
            open(vec, "wb").close()
            open(meta, "w").close()

    def drop_collection(self, collection_name: str) -> None:
        col = self.cols.pop(collection_name, None)
        if col is not None and col.index is not None:
            col.index.close()
        if self.persist:
            for p in self._paths(collection_name):
                if os.path.exists(p):
                    os.remove(p)

    def get_load_state(self, collection_name: str) -> str:
        # The reference compares this value with the string "Loaded" (milvus_service.py:143,154).
        return "Loaded" if self.cols[collection_name].loaded else "NotLoad"

    def load_collection(self, collection_name: str) -> None:
        col = self.cols[collection_name]
        if col.loaded and col.index is not None:
            return
        vecs = self._vectors_from_disk(col) if col.index is None else None
        if col.index is None:
            col.index = VectorIndex(col.dim, device=self.device, capacity=max(len(col.rows), 1024), keep_f32=True)
            if len(col.rows):
                col.index.append(np.ascontiguousarray(vecs), col.levels)
        col.loaded = True

    def release_collection(self, collection_name: str) -> None:
        col = self.cols[collection_name]
        if col.index is not None and self.persist:
            col.index.close()          # frees the HBM table; it is re-read from disk on load
            col.index = None
        col.loaded = False

    def get_collection_stats(self, collection_name: str) -> Dict[str, Any]:
        return {"row_count": len(self.cols[collection_name].rows)}

    def insert(self, collection_name: str, data: Iterable[Dict[str, Any]]) -> Dict[str, Any]:
        col = self.cols[collection_name]
        data = list(data)
        if not data:
            return {"insert_count": 0}
        vecs = np.asarray([d["vector"] for d in data], dtype=np.float32)
        if vecs.ndim != 2 or vecs.shape[1] != col.dim:
            raise ValueError(f"vector dimension mismatch: expected {col.dim}, got {vecs.shape}")
        rows = [{k: v for k, v in d.items() if k != "vector"} for d in data]
        levels = np.asarray([_level_byte(r.get("level", 1)) for r in rows], np.uint8)
        if self.persist:
            _, vec, meta = self._paths(collection_name)
            with open(vec, "ab") as fh:
                fh.write(vecs.astype("<f4").tobytes())
            with open(meta, "a", encoding="utf-8") as fh:
                for r in rows:
                    fh.write(json.dumps(r, ensure_ascii=False) + "\n")
        if col.index is None and (col.loaded or not self.persist):
            col.index = VectorIndex(col.dim, device=self.device, keep_f32=True)
        if col.index is not None:
            col.index.append(vecs, levels)
        first = len(col.rows)
        col.rows.extend(rows)
        col.levels = np.concatenate([col.levels, levels])
        return {"insert_count": len(rows), "ids": list(range(first, first + len(rows)))}

    def insert_device(self, collection_name: str, rows: List[Dict[str, Any]], vecs_dev) -> None:
        """Build path: embeddings already on the GPU (torch float32 [n, dim]); persisted too."""
        col = self.cols[collection_name]
        levels = np.asarray([_level_byte(r.get("level", 1)) for r in rows], np.uint8)
        if self.persist:
            _, vec, meta = self._paths(collection_name)
            with open(vec, "ab") as fh:
                fh.write(vecs_dev.detach().cpu().numpy().astype("<f4").tobytes())
            with open(meta, "a", encoding="utf-8") as fh:
                for r in rows:
                    fh.write(json.dumps(r, ensure_ascii=False) + "\n")
        if col.index is None:
            col.index = VectorIndex(col.dim, device=self.device, keep_f32=True)
        import torch
        col.index.append(vecs_dev.contiguous(), torch.from_numpy(levels).to(vecs_dev.device))
        col.rows.extend(rows)
        col.levels = np.concatenate([col.levels, levels])

    def search(self, collection_name: str, data, limit: int = 10, output_fields: Optional[List[str]] = None,
               **_kw) -> List[List[Hit]]:
        col = self.cols[collection_name]
        if not col.loaded or col.index is None:
            raise N.NativeError(f"collection {collection_name} is not loaded")
        q = np.ascontiguousarray(np.asarray(data, dtype=np.float32))
        if q.ndim == 1:
            q = q[None, :]
        k = max(1, min(int(limit), N.MAX_K))
        _, raw, ids = col.index.search(q, k, weight_mode=N.WEIGHT_NONE)
        fields = output_fields or []
        out = []
        for b in range(q.shape[0]):
            hits = []
            for s, j in zip(raw[b], ids[b]):
                if j < 0:
                    break
                row = col.rows[int(j)]
                hits.append(Hit(id=int(j), distance=float(s), entity={f: row.get(f) for f in fields}))
            out.append(hits)
        return out

    def search_ranked(self, collection_name: str, queries, limit: int):
        """Batched search with the level re-rank done on the GPU (ICD_WEIGHT_RERANK):
        returns (score [B,k], raw [B,k], ids [B,k]) numpy arrays."""
        col = self.cols[collection_name]
        if not col.loaded or col.index is None:
            raise N.NativeError(f"collection {collection_name} is not loaded")
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        return col.index.search(q, max(1, min(int(limit), N.MAX_K)), weight_mode=N.WEIGHT_RERANK)

    def row(self, collection_name: str, row_id: int) -> Dict[str, Any]:
        return self.cols[collection_name].rows[row_id]

    def close(self) -> None:
        for col in self.cols.values():
            if col.index is not None:
                col.index.close()
                col.index = None
            col.loaded = False


def _level_byte(level) -> int:
    try:
        v = int(level)
    except Exception:
        return 0
    return v if 0 <= v <= 255 else 0
