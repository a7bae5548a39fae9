"""Scoring services (host Python) against the reference's own outputs: every case of
tests/golden/scoring_golden.json was produced by the unmodified reference modules."""
import importlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    with open(os.path.join(ROOT, "tests", "golden", "scoring_golden.json"), encoding="utf-8") as fh:
        return json.load(fh)


class _GoldenEmbedding:
    """The recording stub's deterministic 8-d vectors (tests/golden/make_golden.py::_RecordingST)."""

    def encode_query(self, q):
        import hashlib
        h = hashlib.sha256(("query: " + q).encode("utf-8")).digest()
        v = np.frombuffer(h[:32], dtype=np.uint8).astype(np.float32)[:8] - 127.5
        return (v / np.linalg.norm(v)).astype(np.float32)


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.floating,)):
        return float(o)
    if hasattr(o, "__dataclass_fields__"):
        return {k: _jsonable(getattr(o, k)) for k in o.__dataclass_fields__}
    return o


def test_hierarchical_similarity_matches_reference_bit_for_bit():
    g = _golden()
    H = importlib.import_module("rag-project-icd10_b200.services.hierarchical_similarity_service")
    assert len(g["cases"]) == 48
    for case in g["cases"]:
        svc = H.HierarchicalSimilarityService(_GoldenEmbedding() if case["service"] == "with_embedding" else None)
        cands = g["flat"] if case["candidates"] == "flat" else g["nested"]
        ents = g["entities"] if case["entities"] == "ents" else {}
        res = svc.batch_calculate_similarities(case["query"], ents, [dict(c) for c in cands])
        got = [{"record": _jsonable({k: v for k, v in r.items() if k != "similarity_factors"}), "score": s,
                "factors": _jsonable(f)} for r, s, f in res]
        assert len(got) == len(case["result"])
        for a, b in zip(got, case["result"]):
            assert a["record"].get("code") == b["record"].get("code"), (case["query"], case["candidates"])
            if case["service"] == "with_embedding":
                # the stub's cosine goes through float32 vectors: allow 1 ulp-level slack there
                assert abs(a["score"] - b["score"]) < 1e-6
                for k in b["factors"]:
                    assert abs(a["factors"][k] - b["factors"][k]) < 1e-6, k
                assert set(a["record"]) == set(b["record"])
            else:
                assert a == b, (case["query"], case["candidates"], case["entities"])
    svc = H.HierarchicalSimilarityService(None)
    assert {str(k): v for k, v in svc.level_weights.items()} == g["level_weights"]
    f = svc.calculate_enhanced_similarity("急性心肌梗死", g["entities"], dict(g["flat"][3]))[1]
    assert _jsonable(svc.get_similarity_explanation(f)) == g["explanation"]


def test_uncertainty_service_matches_reference():
    g = _golden()
    U = importlib.import_module("rag-project-icd10_b200.services.uncertainty_diagnosis_service")
    unc = U.UncertaintyDiagnosisService()
    for case in g["uncertainty"]:
        assert _jsonable(unc.detect_uncertainty(case["text"])) == case["detect"]
        assert unc.get_uncertainty_explanation(case["text"])["processing_strategy"] == case["explain_strategy"]


def test_update_weights_normalises():
    H = importlib.import_module("rag-project-icd10_b200.services.hierarchical_similarity_service")
    svc = H.HierarchicalSimilarityService(None)
    svc.update_weights({"vector_similarity": 1.0, "nope": 3.0})
    assert abs(sum(svc.factor_weights.values()) - 1.0) < 1e-12 and "nope" not in svc.factor_weights


def test_many_lists_at_once_equals_one_list_at_a_time():
    """the vectorised weighted score over all (diagnosis, candidate) pairs of a request returns
    exactly what batch_calculate_similarities returns per diagnosis (records, scores, factors, order) -- and hence
    the reference's golden outputs."""
    g = _golden()
    H = importlib.import_module("rag-project-icd10_b200.services.hierarchical_similarity_service")
    for service in ("plain", "with_embedding"):
        svc = H.HierarchicalSimilarityService(_GoldenEmbedding() if service == "with_embedding" else None)
        cases = [c for c in g["cases"] if (c["service"] == "with_embedding") == (service == "with_embedding")]
        reqs = [(c["query"], g["entities"] if c["entities"] == "ents" else {},
                 [dict(r) for r in (g["flat"] if c["candidates"] == "flat" else g["nested"])]) for c in cases]
        many = svc.batch_calculate_similarities_many(reqs)
        assert len(many) == len(reqs)
        for (q, ents, cands), got in zip(reqs, many):
            one = svc.batch_calculate_similarities(q, ents, [dict(r) for r in cands])
            assert len(got) == len(one)
            for (ra, sa, fa), (rb, sb, fb) in zip(got, one):
                assert sa == sb and fa == fb and ra.get("code") == rb.get("code")
                assert _jsonable({k: v for k, v in ra.items() if k != "similarity_factors"}) == \
                       _jsonable({k: v for k, v in rb.items() if k != "similarity_factors"})
    assert svc.batch_calculate_similarities_many([]) == []
    F = np.array([[0.97, 0.3, 0.1, 0.99, 0.2, 0.1], [0.5, 0.0, 0.0, 0.3, 0.0, 0.0], [1.0, 1.0, 1.0, 1.0, 1.0, 1.0]])
    scalar = [svc._calculate_weighted_score(H.SimilarityFactors(*row)) for row in F]
    assert svc.weighted_scores(F).tolist() == scalar
