"""HierarchicalSimilarityService -- mirror of
/root/reference/services/hierarchical_similarity_service.py (scoring of <= 100 candidates per
diagnosis).  Python float64 scalar arithmetic kept on the host on purpose:
it is microseconds of work and its string-dependent factors cannot move to the GPU.  Every
formula keeps the reference's operation order so results are bit-identical; pinned against the
reference's own outputs in tests/golden/scoring_golden.json (flat and nested record layouts).

With the dicts MilvusService.search emits (title / metadata nesting, milvus_service.py:297-311)
the keys preferred_zh, level, parent_code and semantic_text are absent at top level, so the
defaults below are what the live path computes -- reproduced, not "fixed".
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Dict, List, Tuple

import numpy as np

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("icd10_b200")

from .uncertainty_diagnosis_service import UncertaintyDiagnosisService


@dataclass
class SimilarityFactors:
    vector_similarity: float = 0.0
    hierarchy_boost: float = 0.0
    entity_match_score: float = 0.0
    semantic_coherence: float = 0.0
    category_alignment: float = 0.0
    context_relevance: float = 0.0

    def __post_init__(self):
        for f in fields(self):
            setattr(self, f.name, float(getattr(self, f.name)))


@dataclass
class HierarchyInfo:
    level: int = 1
    parent_code: str = ""
    category_path: str = ""
    main_category: str = ""
    sub_category: str = ""
    semantic_keywords: List[str] = None

    def __post_init__(self):
        if self.semantic_keywords is None:
            self.semantic_keywords = []


# ICD-10 chapter letter -> (name, keywords, semantic weight); reference :95-142
_CHAPTERS = (
    ("A", "某些传染病和寄生虫病", ("感染", "传染", "病毒", "细菌", "寄生虫", "真菌"), 1.1),
    ("B", "肿瘤", ("癌", "瘤", "肿瘤", "恶性", "良性", "转移"), 1.2),
    ("C", "血液及造血器官疾病", ("血液", "贫血", "白血病", "出血", "凝血"), 1.0),
    ("E", "内分泌、营养和代谢疾病", ("糖尿病", "甲状腺", "代谢", "内分泌", "营养"), 1.1),
    ("I", "循环系统疾病", ("心脏", "血管", "高血压", "心肌", "循环"), 1.2),
    ("J", "呼吸系统疾病", ("肺", "呼吸", "咳嗽", "气管", "支气管"), 1.1),
    ("K", "消化系统疾病", ("胃", "肠", "肝", "消化", "腹泻"), 1.0),
    ("N", "泌尿生殖系统疾病", ("肾", "膀胱", "泌尿", "生殖", "尿"), 1.0),
    ("S", "损伤、中毒和外因的某些其他后果", ("损伤", "外伤", "骨折", "中毒", "烧伤"), 0.9),
)

_LEVEL_BOOST = {1: 0.15, 2: 0.20, 3: 0.10}  # reference :285-291

# divisors the reference divides the configured weights by (they cancel at default weights);
# reference :491-506
_NORMALISERS = {"hierarchy_boost": 0.2, "entity_match_score": 0.15, "semantic_coherence": 0.08,
                "category_alignment": 0.04, "context_relevance": 0.03}

_DESCRIPTIONS = {
    "vector_similarity": "基础向量相似度", "hierarchy_boost": "ICD-10层级增强分数",
    "entity_match_score": "医学实体匹配分数", "semantic_coherence": "语义一致性分数",
    "category_alignment": "ICD类别对齐分数", "context_relevance": "上下文相关性分数",
}


def _cosine(u, v) -> float:
    """sklearn.metrics.pairwise.cosine_similarity([u], [v])[0][0] (reference :237,408)."""
    u = np.asarray(u, dtype=np.float64).ravel()
    v = np.asarray(v, dtype=np.float64).ravel()
    nu, nv = np.linalg.norm(u), np.linalg.norm(v)
    nu = nu if nu != 0.0 else 1.0
    nv = nv if nv != 0.0 else 1.0
    return float(np.dot(u / nu, v / nv))


class HierarchicalSimilarityService:
    def __init__(self, embedding_service=None, ner_service=None):
        self.embedding_service = embedding_service
        self.ner_service = ner_service
        self.uncertainty_service = UncertaintyDiagnosisService()
        self.level_weights = {1: 1.2, 2: 1.0, 3: 0.8}  # reference :69-73
        self.factor_weights = {
            "vector_similarity": 0.50, "hierarchy_boost": 0.20, "entity_match_score": 0.15,
            "semantic_coherence": 0.08, "category_alignment": 0.04, "context_relevance": 0.03,
        }
        self.main_categories = self._load_main_categories()
        self.similarity_cache = {}
        logger.info(f"层级相似度服务初始化完成，权重配置: {self.factor_weights}")

    def _load_main_categories(self) -> Dict[str, Dict[str, Any]]:
        return {letter: {"name": name, "keywords": list(kws), "semantic_weight": w}
                for letter, name, kws, w in _CHAPTERS}

    # ------------------------------------------------------------------ one candidate
    def calculate_enhanced_similarity(self, query_text: str, query_entities: Dict[str, List[Dict]],
                                      candidate_record: Dict[str, Any]) -> Tuple[float, SimilarityFactors]:
        factors = SimilarityFactors()
        try:
            title = candidate_record.get("preferred_zh", "").strip()
            wanted = query_text.strip()
            exact = title == wanted                                        # reference :162-165
            factors.vector_similarity = self._calculate_vector_similarity(query_text, candidate_record)
            if exact and factors.vector_similarity < 0.9:                 # reference :172-176
                factors.vector_similarity = 1.0
            factors.hierarchy_boost = self._calculate_hierarchy_boost(query_text, query_entities, candidate_record)
            factors.entity_match_score = self._calculate_entity_match_score(query_entities, candidate_record)
            factors.semantic_coherence = self._calculate_semantic_coherence(query_text, candidate_record)
            factors.category_alignment = self._calculate_category_alignment(query_entities, candidate_record)
            factors.context_relevance = self._calculate_context_relevance(query_text, candidate_record)
            total = self._calculate_weighted_score(factors)
            if exact:
                total = max(total, 1.5)                                   # reference :206-208
            return float(total), factors
        except Exception as e:  # same degrade contract as the reference (:215-219)
            logger.error(f"增强相似度计算失败: {e}")
            return float(candidate_record.get("score", 0.0)), factors

    # reference :221-242
    def _calculate_vector_similarity(self, query_text: str, candidate_record: Dict[str, Any]) -> float:
        try:
            if not self.embedding_service:
                return candidate_record.get("score", 0.0)
            if "score" in candidate_record:
                return float(candidate_record["score"])
            qv = self.embedding_service.encode_query(query_text)
            text = candidate_record.get("semantic_text", candidate_record.get("preferred_zh", ""))
            cv = self.embedding_service.encode_query(text)
            return float(max(_cosine(qv, cv), 0.0))
        except Exception as e:
            logger.warning(f"向量相似度计算失败: {e}")
            return candidate_record.get("score", 0.0)

    # reference :244-283
    def _calculate_hierarchy_boost(self, query_text, query_entities, candidate_record) -> float:
        try:
            level = candidate_record.get("level", 1)
            code = candidate_record.get("code", "")
            parent = candidate_record.get("parent_code", "")
            total = 0.0
            total += self._get_level_boost_factor(level) * 0.3
            chapter = code[0] if code else ""
            if chapter in self.main_categories:
                total += self._calculate_category_semantic_boost(query_text, query_entities,
                                                                 self.main_categories[chapter]) * 0.4
            if parent:
                total += self._calculate_parent_child_boost(query_entities, code, parent) * 0.3
            return float(min(total, 0.3))
        except Exception as e:
            logger.warning(f"层级增强分数计算失败: {e}")
            return 0.0

    def _get_level_boost_factor(self, level: int) -> float:
        return float(_LEVEL_BOOST.get(level, 0.10))

    # reference :293-328
    def _calculate_category_semantic_boost(self, query_text, query_entities, category_info) -> float:
        try:
            kws = category_info.get("keywords", [])
            weight = category_info.get("semantic_weight", 1.0)
            total = 0.0
            lowered = query_text.lower()
            hit = sum(1 for kw in kws if kw in lowered)
            if hit > 0:
                total += ((hit / len(kws)) * 0.3) * weight
            for ent in query_entities.get("disease", []):
                text = ent.get("text", "").lower()
                n = sum(1 for kw in kws if kw in text)
                if n > 0:
                    total += ((n / len(kws)) * 0.2) * ent.get("confidence", 0.5)
            return float(min(total, 0.4))
        except Exception as e:
            logger.warning(f"类别语义增强计算失败: {e}")
            return 0.0

    # reference :330-339
    def _calculate_parent_child_boost(self, query_entities, code: str, parent_code: str) -> float:
        return 0.1 if (len(code) > len(parent_code) and code.startswith(parent_code)) else 0.0

    # reference :341-386
    def _calculate_entity_match_score(self, query_entities, candidate_record) -> float:
        try:
            hay = f"{candidate_record.get('preferred_zh', '').lower()} {candidate_record.get('semantic_text', '').lower()}"
            total = 0.0
            for ent in query_entities.get("disease", []):
                text, conf = ent.get("text", "").lower(), ent.get("confidence", 0.5)
                if text in hay:
                    total += conf * 0.4
                elif any(word in hay for word in text.split()):
                    total += conf * 0.2
            for kind, gain in (("symptom", 0.2), ("anatomy", 0.1)):
                for ent in query_entities.get(kind, []):
                    if ent.get("text", "").lower() in hay:
                        total += ent.get("confidence", 0.5) * gain
            return float(min(total, 1.0))
        except Exception as e:
            logger.warning(f"实体匹配分数计算失败: {e}")
            return 0.0

    # reference :388-412
    def _calculate_semantic_coherence(self, query_text, candidate_record) -> float:
        try:
            if not self.embedding_service:
                return 0.5
            text = candidate_record.get("semantic_text", "")
            if not text:
                return 0.3
            qv = self.embedding_service.encode_query(query_text)
            sv = self.embedding_service.encode_query(text)
            return max(_cosine(qv, sv), 0.0)
        except Exception as e:
            logger.warning(f"语义一致性计算失败: {e}")
            return 0.5

    # reference :414-449
    def _calculate_category_alignment(self, query_entities, candidate_record) -> float:
        try:
            code = candidate_record.get("code", "")
            if not code or code[0] not in self.main_categories:
                return 0.0
            kws = self.main_categories[code[0]].get("keywords", [])
            aligned, count = 0.0, 0
            for _kind, ents in query_entities.items():
                for ent in ents:
                    count += 1
                    text = ent.get("text", "").lower()
                    if any(kw in text for kw in kws):
                        aligned += ent.get("confidence", 0.5)
            return float(aligned / count) if count > 0 else 0.0
        except Exception as e:
            logger.warning(f"类别对齐度计算失败: {e}")
            return 0.0

    # reference :451-473
    def _calculate_context_relevance(self, query_text, candidate_record) -> float:
        try:
            title = candidate_record.get("preferred_zh", "")
            lq, lt = len(query_text), len(title)
            length_sim = 1.0 - abs(lq - lt) / max(lq, lt, 1)
            a, b = set(query_text), set(title)
            union = a | b
            overlap = len(a & b) / len(union) if union else 0
            return max(length_sim * 0.3 + overlap * 0.7, 0.0)
        except Exception as e:
            logger.warning(f"上下文相关性计算失败: {e}")
            return 0.5

    # reference :475-518: additive enhancement on top of the vector score, capped at 1.8
    def _calculate_weighted_score(self, factors: SimilarityFactors) -> float:
        try:
            w = self.factor_weights
            base = factors.vector_similarity
            precise = base > 0.95
            extra = 0.0
            extra += factors.hierarchy_boost * w["hierarchy_boost"] / _NORMALISERS["hierarchy_boost"] * (0.5 if precise else 1.0)
            extra += factors.entity_match_score * w["entity_match_score"] / _NORMALISERS["entity_match_score"]
            if factors.semantic_coherence > base:
                extra += (factors.semantic_coherence - base) * w["semantic_coherence"] / _NORMALISERS["semantic_coherence"]
            extra += factors.category_alignment * w["category_alignment"] / _NORMALISERS["category_alignment"]
            extra += factors.context_relevance * w["context_relevance"] / _NORMALISERS["context_relevance"]
            if precise:
                extra += 0.15
            return float(min(base + extra, 1.8))
        except Exception as e:
            logger.error(f"加权分数计算失败: {e}")
            return float(factors.vector_similarity)

    # ------------------------------------------------------------------ a candidate list
    # reference :520-579
    def batch_calculate_similarities(self, query_text: str, query_entities: Dict[str, List[Dict]],
                                     candidate_records: List[Dict[str, Any]]
                                     ) -> List[Tuple[Dict[str, Any], float, SimilarityFactors]]:
        logger.info(f"开始批量计算 {len(candidate_records)} 个候选记录的增强相似度")
        core, candidates = self.uncertainty_service.process_uncertainty_query(query_text, candidate_records)
        if core != query_text:
            logger.info(f"不确定性处理: '{query_text}' -> '{core}'")
        out = []
        for rec in candidates:
            try:
                score, factors = self.calculate_enhanced_similarity(core, query_entities, rec)
                item = rec.copy()
                item["enhanced_score"] = score
                item["original_score"] = rec.get("original_score", rec.get("score", 0.0))
                item["similarity_factors"] = factors
                if "uncertainty_boost" in rec:
                    item["uncertainty_boost"] = rec["uncertainty_boost"]
                    item["uncertainty_weight"] = rec["uncertainty_weight"]
                out.append((item, score, factors))
            except Exception as e:
                logger.error(f"记录 {rec.get('code', 'unknown')} 的相似度计算失败: {e}")
                out.append((rec, rec.get("score", 0.0), SimilarityFactors()))
        out.sort(key=lambda t: t[1], reverse=True)
        logger.info(f"批量相似度计算完成，平均增强分数: {float(np.mean([t[1] for t in out])):.4f}")
        return out

    # ------------------------------------------------------------------ many candidate lists at once
    def weighted_scores(self, F: np.ndarray) -> np.ndarray:
        """_calculate_weighted_score over a [P, 6] float64 array of factors (columns in SimilarityFactors order), one
        vectorised pass; the operation order per element is the scalar function's, so results are bit-identical."""
        w, n = self.factor_weights, _NORMALISERS
        base, hb, em, sc, ca, cr = (F[:, i] for i in range(6))
        precise = base > 0.95
        extra = np.zeros_like(base)
        extra = extra + hb * w["hierarchy_boost"] / n["hierarchy_boost"] * np.where(precise, 0.5, 1.0)
        extra = extra + em * w["entity_match_score"] / n["entity_match_score"]
        extra = extra + np.where(sc > base, (sc - base) * w["semantic_coherence"] / n["semantic_coherence"], 0.0)
        extra = extra + ca * w["category_alignment"] / n["category_alignment"]
        extra = extra + cr * w["context_relevance"] / n["context_relevance"]
        extra = extra + np.where(precise, 0.15, 0.0)
        return np.minimum(base + extra, 1.8)

    def batch_calculate_similarities_many(self, requests):
        """All diagnoses of a request in one call: `requests` is a list of (query_text, query_entities,
        candidate_records) -- e.g. one entry per extracted diagnosis with the rows MilvusService.search_batch
        returned for it.

        The reference's per-candidate loop encodes TWO texts per candidate for the semantic-coherence factor
        (_calculate_semantic_coherence, reference :388-412: the query and the candidate's semantic_text, batch 1
        each) -- 40 encoder forwards for the 20 candidates of one diagnosis, the dominant cost of a request.  Here
        every distinct text of the whole request goes through the encoder ONCE, in one batch
        (EmbeddingService.encode_queries), and the factor is a dot product of cached vectors; the other
        string-dependent factors are computed per (query, candidate) on the host exactly as in
        batch_calculate_similarities, and the weighted score of ALL pairs is one vectorised pass (weighted_scores).
        Returns one list per request with the records, factors and order of batch_calculate_similarities(query_text,
        query_entities, candidate_records) on each -- bit-identical when the embedding service is deterministic per
        text (the golden test), within the encoder's batch-composition noise (~1e-6) on the GPU."""
        plans, rows = [], []          # rows: (request index, record, factors | None, exact)
        prepared = [(self.uncertainty_service.process_uncertainty_query(q, c), e) for q, e, c in requests]
        vec_of = self._encode_request_texts(prepared)
        for ri, ((core, candidates), query_entities) in enumerate(prepared):
            plans.append(len(candidates))
            for rec in candidates:
                factors = SimilarityFactors()
                try:
                    exact = rec.get("preferred_zh", "").strip() == core.strip()
                    factors.vector_similarity = self._calculate_vector_similarity(core, rec)
                    if exact and factors.vector_similarity < 0.9:
                        factors.vector_similarity = 1.0
                    factors.hierarchy_boost = self._calculate_hierarchy_boost(core, query_entities, rec)
                    factors.entity_match_score = self._calculate_entity_match_score(query_entities, rec)
                    factors.semantic_coherence = self._semantic_coherence_cached(core, rec, vec_of)
                    factors.category_alignment = self._calculate_category_alignment(query_entities, rec)
                    factors.context_relevance = self._calculate_context_relevance(core, rec)
                    rows.append((ri, rec, factors, exact, True))
                except Exception as e:   # the scalar path's degrade contract: the candidate keeps its own score
                    logger.error(f"增强相似度计算失败: {e}")
                    rows.append((ri, rec, factors, False, False))
        F = np.array([[getattr(f, fld.name) for fld in fields(SimilarityFactors)] for _, _, f, _, _ in rows],
                     dtype=np.float64).reshape(len(rows), 6)
        totals = self.weighted_scores(F) if len(rows) else np.zeros(0)
        out = [[] for _ in requests]
        for (ri, rec, factors, exact, ok), total in zip(rows, totals):
            if ok:
                score = float(max(float(total), 1.5)) if exact else float(total)
            else:
                score = float(rec.get("score", 0.0))
            item = rec.copy()
            item["enhanced_score"] = score
            item["original_score"] = rec.get("original_score", rec.get("score", 0.0))
            item["similarity_factors"] = factors
            if "uncertainty_boost" in rec:
                item["uncertainty_boost"] = rec["uncertainty_boost"]
                item["uncertainty_weight"] = rec["uncertainty_weight"]
            out[ri].append((item, score, factors))
        for lst in out:
            lst.sort(key=lambda t: t[1], reverse=True)
        return out

    def _encode_request_texts(self, prepared) -> Dict[str, Any]:
        """text -> embedding for every distinct text the semantic-coherence factor of a request needs (the core query of
        every diagnosis, the semantic_text of every candidate): one encode_queries call, or {} when that fails / there
        is no embedding service (the factor then falls back to the per-candidate path and its defaults)."""
        if not self.embedding_service:
            return {}
        texts, seen = [], set()
        for (core, candidates), _entities in prepared:
            for t in [core] + [rec.get("semantic_text", "") for rec in candidates]:
                if t and t not in seen:
                    seen.add(t)
                    texts.append(t)
        if not texts:
            return {}
        try:
            many = getattr(self.embedding_service, "encode_queries", None)
            vecs = many(texts) if many is not None else [self.embedding_service.encode_query(t) for t in texts]
            return {t: v for t, v in zip(texts, vecs)}
        except Exception as e:
            logger.warning(f"批量语义向量计算失败: {e}")
            return {}

    def _semantic_coherence_cached(self, query_text, candidate_record, vec_of) -> float:
        """_calculate_semantic_coherence (reference :388-412) over the vectors of _encode_request_texts."""
        if not self.embedding_service:
            return 0.5
        text = candidate_record.get("semantic_text", "")
        if not text:
            return 0.3
        qv, sv = vec_of.get(query_text), vec_of.get(text)
        if qv is None or sv is None:
            return self._calculate_semantic_coherence(query_text, candidate_record)
        try:
            return max(_cosine(qv, sv), 0.0)
        except Exception as e:
            logger.warning(f"语义一致性计算失败: {e}")
            return 0.5

    # reference :581-624
    def get_similarity_explanation(self, factors: SimilarityFactors) -> Dict[str, Any]:
        detail = {}
        for name, desc in _DESCRIPTIONS.items():
            value = getattr(factors, name)
            detail[name] = {"score": value, "weight": self.factor_weights[name],
                            "contribution": value * self.factor_weights[name], "description": desc}
        return {"total_score": self._calculate_weighted_score(factors), "factors": detail}

    # reference :626-638
    def update_weights(self, new_weights: Dict[str, float]):
        for name, value in new_weights.items():
            if name in self.factor_weights:
                self.factor_weights[name] = value
                logger.info(f"权重更新: {name} = {value}")
        total = sum(self.factor_weights.values())
        if total != 1.0:
            logger.warning(f"权重总和不为1.0: {total}，自动归一化")
            for name in self.factor_weights:
                self.factor_weights[name] /= total
