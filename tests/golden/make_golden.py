#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the REFERENCE's own Python in the build container.

Run here only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference's two third-party engines (sentence_transformers, pymilvus) are not
installable offline, so they are replaced by recording stubs *around* the unmodified
reference classes: what is pinned is the reference's own code (text preparation, CSV rules,
hit layout / level re-rank / sort, scoring formulas), not the engines' arithmetic.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


# ------------------------------------------------------------------ engine stubs
class _RecordingST:
    """Stands in for sentence_transformers.SentenceTransformer; records what it is asked."""
    calls = []

    def __init__(self, name, device=None):
        self.name, self.device, self.max_seq_length = name, device, 128

    def get_sentence_embedding_dimension(self):
        return 8

    @staticmethod
    def _vec(text):
        h = hashlib.sha256(text.encode("utf-8")).digest()
        v = np.frombuffer(h[:32], dtype=np.uint8).astype(np.float32)[:8] - 127.5
        return (v / np.linalg.norm(v)).astype(np.float32)

    def encode(self, texts, **kw):
        _RecordingST.calls.append({"texts": texts, "kwargs": {k: kw[k] for k in sorted(kw)}})
        if isinstance(texts, str):
            return self._vec(texts)
        return np.stack([self._vec(t) for t in texts])


class _Hit(dict):
    """pymilvus Hit: .get falls through to the entity fields."""

    def get(self, key, default=None):
        if key in self:
            return dict.get(self, key)
        return dict.get(self, "entity", {}).get(key, default)


class _Schema:
    def __init__(self, **kw):
        self.kw, self.fields = kw, []

    def add_field(self, **kw):
        self.fields.append({k: str(v) for k, v in kw.items()})


class _IndexParams:
    def __init__(self):
        self.indexes = []

    def add_index(self, **kw):
        self.indexes.append(kw)


class _ExactClient:
    """Stands in for pymilvus.MilvusClient: FLAT/IP == exact fp32 inner product."""
    log = []

    def __init__(self, **kw):
        self.kw, self.cols = kw, {}

    def has_collection(self, collection_name):
        return collection_name in self.cols

    def create_schema(self, **kw):
        return _Schema(**kw)

    def prepare_index_params(self):
        return _IndexParams()

    def create_collection(self, collection_name, schema, index_params):
        _ExactClient.log.append({"create": collection_name, "schema_kw": schema.kw,
                                 "fields": schema.fields, "indexes": index_params.indexes})
        self.cols[collection_name] = []

    def get_load_state(self, collection_name):
        return {"state": "<LoadState: Loaded>"}

    def load_collection(self, collection_name):
        pass

    def release_collection(self, collection_name):
        pass

    def drop_collection(self, collection_name):
        self.cols.pop(collection_name, None)

    def get_collection_stats(self, collection_name):
        return {"row_count": len(self.cols[collection_name])}

    def insert(self, collection_name, data):
        self.cols[collection_name].extend(data)

    def search(self, collection_name, data, limit, output_fields):
        rows = self.cols[collection_name]
        mat = np.asarray([r["vector"] for r in rows], np.float32)
        out = []
        for q in data:
            s = mat @ np.asarray(q, np.float32)
            order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))[:limit]
            out.append([_Hit(id=int(j), distance=float(s[j]),
                             entity={f: rows[j][f] for f in output_fields}) for j in order])
        return out

    def close(self):
        pass


def _install_stubs():
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = _RecordingST
    sys.modules["sentence_transformers"] = st
    pm = types.ModuleType("pymilvus")
    pm.MilvusClient = _ExactClient

    class DataType:
        INT64, FLOAT_VECTOR, VARCHAR, BOOL, INT32 = "INT64", "FLOAT_VECTOR", "VARCHAR", "BOOL", "INT32"
    pm.DataType = DataType
    sys.modules["pymilvus"] = pm


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, np.ndarray):
        return o.tolist()
    if hasattr(o, "__dataclass_fields__"):
        return {k: _jsonable(getattr(o, k)) for k in o.__dataclass_fields__}
    return o


def main():
    os.environ["EMBEDDING_MODEL_NAME"] = "shibing624/text2vec-base-chinese"
    os.environ["EMBEDDING_DEVICE"] = "cpu"
    os.environ["MILVUS_DB_PATH"] = "/tmp/golden_db/milvus.db"
    os.environ["MILVUS_COLLECTION_NAME"] = "icd10_golden"
    _install_stubs()
    sys.path.insert(0, REF)
    os.chdir("/tmp")
    from loguru import logger
    logger.remove()

    # ---- 1. CSV -> records, by the reference's DatabaseBuilder (tools/build_database.py:62-192)
    from tools.build_database import DatabaseBuilder
    builder = DatabaseBuilder()
    records = builder.load_csv_data(os.path.join(REF, "data/ICD_10v601.csv"))
    canon = json.dumps(records, ensure_ascii=False, sort_keys=True).encode("utf-8")
    wanted = {"A00", "A00.0", "A00.001", "A00.901", "A01.003+G01*", "A01.005+J17.0*", "M998902/3"}
    rec_golden = {
        "count": len(records),
        "level_counts": {str(l): sum(1 for r in records if r["level"] == l) for l in (1, 2, 3)},
        "sha256": hashlib.sha256(canon).hexdigest(),
        "samples": [r for r in records if r["code"] in wanted] + records[20000:20010] + records[-5:],
        "batch_sizes": {str(n): builder._calculate_optimal_batch_size(n) for n in (1, 999, 1000, 9999, 10000, 40474, 49999, 50000)},
    }
    with open(os.path.join(HERE, "records_golden.json"), "w", encoding="utf-8") as fh:
        json.dump(rec_golden, fh, ensure_ascii=False, indent=1)

    # ---- 2. EmbeddingService text preparation (services/embedding_service.py:68-120)
    from services.embedding_service import EmbeddingService
    es = EmbeddingService()
    _RecordingST.calls.clear()
    probe = ["急性胃肠炎", "query: 已带前缀", "passage: 已带前缀", "", " 前导空格", "Query: 大写不算"]
    outs = {}
    for t in probe:
        es.encode_single(t)
        es.encode_query(t)
    outs["encode_batch_empty"] = es.encode_batch([])
    eb = es.encode_batch(probe, show_progress=False)
    outs["encode_batch_type"] = [type(eb).__name__, type(eb[0]).__name__, type(eb[0][0]).__name__]
    es.encode_icd_record({"code": "A00", "preferred_zh": "霍乱"})
    es.encode_icd_record({"code": "A00", "preferred_zh": "  "})
    es.encode_icd_record({"preferred_zh": ""})
    info = es.get_model_info()
    te = es.test_embedding("测试")
    emb_golden = {"calls": _jsonable(_RecordingST.calls), "outs": outs, "model_info": _jsonable(info),
                  "test_embedding_keys": sorted(te.keys()), "test_embedding_shape": list(te["embedding_shape"])}
    with open(os.path.join(HERE, "embedding_service_golden.json"), "w", encoding="utf-8") as fh:
        json.dump(emb_golden, fh, ensure_ascii=False, indent=1)

    # ---- 3. MilvusService over an exact-IP client (services/milvus_service.py)
    from services.milvus_service import MilvusService
    ms = MilvusService(embedding_service=es)
    rng = np.random.default_rng(20261017)
    sub = records[:120] + records[5000:5080] + [r for r in records if r["code"] in wanted]
    vecs = rng.standard_normal((len(sub), 8)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    ok = ms.insert_records(sub, [v for v in vecs])
    queries = rng.standard_normal((6, 8)).astype(np.float32)
    queries /= np.linalg.norm(queries, axis=1, keepdims=True)
    searches = []
    for q in queries:
        for k in (1, 5, 10):
            searches.append({"query": q.tolist(), "top_k": k, "hits": _jsonable(ms.search(q, top_k=k))})
    try:
        ms.insert_records(sub[:2], [vecs[0]])
        mismatch = "no error"
    except ValueError as e:
        mismatch = "ValueError: " + str(e)
    milvus_golden = {
        "create_log": _jsonable(_ExactClient.log),
        "dimension": ms.dimension,
        "insert_ok": ok,
        "codes": [r["code"] for r in sub],
        "vectors": vecs.tolist(),
        "searches": searches,
        "length_mismatch": mismatch,
        "stats": _jsonable(ms.get_collection_stats()),
        "memory_usage": _jsonable(ms.get_memory_usage()),
        "load_state": _jsonable(ms.get_collection_load_state()),
        "test_connection": _jsonable(ms.test_connection()),
        "release": _jsonable(ms.release_collection()),
        "health_keys": sorted(ms.health_check().keys()),
        "level_weights": {str(l): ms._calculate_level_weight(l) for l in (0, 1, 2, 3, 4)},
    }
    hits_for_scoring = ms.search(queries[0], top_k=10)
    with open(os.path.join(HERE, "milvus_service_golden.json"), "w", encoding="utf-8") as fh:
        json.dump(milvus_golden, fh, ensure_ascii=False, indent=1)

    # ---- 4. scoring services (hierarchical_similarity_service.py, uncertainty_diagnosis_service.py)
    from services.hierarchical_similarity_service import HierarchicalSimilarityService
    from services.uncertainty_diagnosis_service import UncertaintyDiagnosisService
    flat = [
        {"code": "I21.9", "preferred_zh": "急性心肌梗死，未特指", "level": 3, "parent_code": "I21",
         "category_path": "I > I21 > I21.9", "semantic_text": "急性心肌梗死 | 循环系统疾病 | ICD-10: I21.9", "score": 0.85},
        {"code": "I47.9", "preferred_zh": "阵发性心动过速，未特指", "level": 3, "parent_code": "I47",
         "category_path": "I > I47 > I47.9", "semantic_text": "阵发性心动过速 | 心律失常 | ICD-10: I47.9", "score": 0.72},
        {"code": "I25.9", "preferred_zh": "慢性缺血性心脏病，未特指", "level": 3, "parent_code": "I25",
         "category_path": "I > I25 > I25.9", "semantic_text": "慢性缺血性心脏病 | 循环系统疾病 | ICD-10: I25.9", "score": 0.68},
        {"code": "I21", "preferred_zh": "急性心肌梗死", "level": 1, "parent_code": "",
         "category_path": "I21", "semantic_text": "急性心肌梗死 | ICD-10: I21", "score": 0.97},
        {"code": "J18.901", "preferred_zh": "肺炎", "level": 3, "parent_code": "J18.9",
         "category_path": "J18 > J18.9 > J18.901", "semantic_text": "肺炎 | 未特指的肺炎 | ICD-10: J18.901", "score": 0.91},
        {"code": "J18.9", "preferred_zh": "未特指的肺炎", "level": 2, "parent_code": "J18",
         "category_path": "J18 > J18.9", "semantic_text": "未特指的肺炎 | 肺炎 | ICD-10: J18.9", "score": 0.88},
        {"code": "Z99", "preferred_zh": "其他肺炎", "level": 1, "score": 0.5},
        {"code": "", "preferred_zh": "无编码", "score": 0.4},
    ]
    ents = {
        "disease": [{"text": "急性心肌梗死", "confidence": 0.95, "start": 0, "end": 6},
                    {"text": "心律失常", "confidence": 0.88, "start": 7, "end": 11},
                    {"text": "肺炎 感染"}],
        "anatomy": [{"text": "心肌", "confidence": 0.85, "start": 2, "end": 4}],
        "symptom": [{"text": "咳嗽", "confidence": 0.7}],
    }
    cases = []
    for label, emb in (("no_embedding", None), ("with_embedding", es)):
        svc = HierarchicalSimilarityService(emb)
        for q in ("急性心肌梗死伴心律失常", "急性心肌梗死", "肺炎待查", "疑似肺炎？", "肺炎", "胃肠 感染 咳嗽"):
            for cname, cands in (("flat", flat), ("nested", hits_for_scoring)):
                for ename, e in (("ents", ents), ("noents", {})):
                    res = svc.batch_calculate_similarities(q, e, [dict(c) for c in cands])
                    cases.append({"service": label, "query": q, "candidates": cname, "entities": ename,
                                  "result": [{"record": _jsonable({k: v for k, v in r.items() if k != "similarity_factors"}),
                                              "score": s, "factors": _jsonable(f)} for r, s, f in res]})
    unc = UncertaintyDiagnosisService()
    unc_cases = [{"text": t, "detect": _jsonable(unc.detect_uncertainty(t)),
                  "explain_strategy": unc.get_uncertainty_explanation(t)["processing_strategy"]}
                 for t in ("肺炎待查", "疑似肺炎？", "糖尿病?可能", "高血压", "，发热原因不明。", "不能排除结核 考虑肿瘤")]
    svc = HierarchicalSimilarityService(None)
    expl = svc.get_similarity_explanation(cases[0]["result"][0] and
                                          svc.calculate_enhanced_similarity("急性心肌梗死", ents, dict(flat[3]))[1])
    with open(os.path.join(HERE, "scoring_golden.json"), "w", encoding="utf-8") as fh:
        json.dump({"flat": flat, "nested": _jsonable(hits_for_scoring), "entities": ents, "cases": cases,
                   "uncertainty": unc_cases, "explanation": _jsonable(expl),
                   "level_weights": {str(k): v for k, v in svc.level_weights.items()}},
                  fh, ensure_ascii=False, indent=1)
    print("records", rec_golden["count"], rec_golden["sha256"][:16], "| cases", len(cases),
          "| searches", len(searches))


if __name__ == "__main__":
    main()
