"""UncertaintyDiagnosisService -- mirror of /root/reference/services/uncertainty_diagnosis_service.py.

Runs unchanged *inside* HierarchicalSimilarityService.batch_calculate_similarities
(reference hierarchical_similarity_service.py:540-542), so its effect on candidate scores and
order is part of the hot path's observable behaviour.  Host-side string rules;
pinned against the reference's own outputs in tests/golden/scoring_golden.json.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List, Tuple

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("icd10_b200")

# (type, weight, description, markers) in the reference's evaluation order
# reference: uncertainty_diagnosis_service.py:19-40
_UNCERTAINTY_CLASSES = (
    ("explicit_uncertainty", 1.0, "明确不确定性", ("待查", "待诊", "待确诊", "待定", "排除", "？", "?")),
    ("suspected", 0.9, "疑似性", ("疑似", "疑为", "考虑", "可能", "拟诊", "倾向")),
    ("degree_uncertainty", 0.8, "程度不确定性", ("不除外", "不能排除", "不明原因", "原因不明", "性质待定")),
)

# "unspecified" title rules, first hit wins; reference: :43-70
_EXACT_TEMPLATES = ("未特指的{}", "{}，未特指", "{}未特指")
_OTHER_TEMPLATES = ("其他{}", "{}，其他", "不明{}", "{}不明")
_CODE_DOT9 = re.compile(r"\.9\d*$")
_EDGE_PUNCT = re.compile(r"^[，。、\s]+|[，。、\s]+$")


class UncertaintyDiagnosisService:
    def __init__(self):
        self.uncertainty_patterns = {
            name: {"patterns": list(marks), "weight": w, "description": desc}
            for name, w, desc, marks in _UNCERTAINTY_CLASSES
        }
        self.icd_unspecified_patterns = {
            "exact_unspecified": {"patterns": list(_EXACT_TEMPLATES), "boost": 0.3, "description": "精确未特指匹配"},
            "contains_unspecified": {"patterns": ["未特指"], "boost": 0.25, "description": "包含未特指"},
            "other_uncertainty": {"patterns": list(_OTHER_TEMPLATES), "boost": 0.2, "description": "其他不确定性"},
            "code_structure": {"code_pattern": _CODE_DOT9.pattern, "boost": 0.15, "description": "编码结构暗示"},
        }
        logger.info("不确定性诊断处理服务初始化完成")

    # reference :74-125
    def detect_uncertainty(self, text: str) -> Dict[str, Any]:
        lowered = text.lower()
        hits: List[Dict[str, Any]] = []
        kind, weight = None, 0.0
        for name, cfg in self.uncertainty_patterns.items():
            for mark in cfg["patterns"]:
                at = lowered.find(mark.lower())
                if at < 0:
                    continue
                kind = name                      # last matching class wins, as in the reference
                weight = max(weight, cfg["weight"])
                hits.append({"pattern": mark, "type": name, "weight": cfg["weight"], "position": at})
        cleaned = text
        if hits:
            for h in hits:
                cleaned = re.sub(re.escape(h["pattern"]), "", cleaned, flags=re.IGNORECASE)
            cleaned = re.sub(r"\s+", " ", cleaned).strip()
            cleaned = _EDGE_PUNCT.sub("", cleaned)
        result = {
            "has_uncertainty": bool(hits),
            "uncertainty_type": kind,
            "uncertainty_weight": weight,
            "matched_patterns": [h["pattern"] for h in hits],
            "clean_text": cleaned,
            "uncertainty_indicators": hits,
        }
        logger.debug(f"不确定性检测: '{text}' -> {result}")
        return result

    # reference :127-188
    def calculate_unspecified_boost(self, candidate_record: Dict[str, Any], clean_diagnosis: str) -> float:
        title = candidate_record.get("preferred_zh", "").lower()
        code = candidate_record.get("code", "")
        core = clean_diagnosis.lower()
        rules = self.icd_unspecified_patterns
        boost, why = 0.0, None
        if any(t.format(core) in title for t in rules["exact_unspecified"]["patterns"]):
            boost, why = rules["exact_unspecified"]["boost"], "exact_unspecified"
        elif any(p in title for p in rules["contains_unspecified"]["patterns"]):
            boost, why = rules["contains_unspecified"]["boost"], "contains_unspecified"
        elif any(t.format(core) in title for t in rules["other_uncertainty"]["patterns"]):
            boost, why = rules["other_uncertainty"]["boost"], "other_uncertainty"
        elif re.search(rules["code_structure"]["code_pattern"], code):
            boost, why = rules["code_structure"]["boost"], "code_structure"
        if boost > 0:
            logger.info(f"未特指加权: '{clean_diagnosis}' -> '{title}' (+{boost:.3f}, 类型: {[why]})")
        return boost

    # reference :190-238
    def process_uncertainty_query(self, query_text: str,
                                  candidate_records: List[Dict[str, Any]]) -> Tuple[str, List[Dict[str, Any]]]:
        found = self.detect_uncertainty(query_text)
        if not found["has_uncertainty"]:
            return query_text, candidate_records
        core, weight = found["clean_text"], found["uncertainty_weight"]
        logger.info(f"检测到不确定性诊断: '{query_text}' -> '{core}' (权重: {weight})")
        out = []
        for rec in candidate_records:
            item = rec.copy()
            boost = self.calculate_unspecified_boost(rec, core)
            if boost > 0:
                before = item.get("score", 0.0)
                item["score"] = before + (boost * weight)
                item["uncertainty_boost"] = boost
                item["uncertainty_weight"] = weight
                item["original_score"] = before
            out.append(item)
        out.sort(key=lambda r: r.get("score", 0.0), reverse=True)
        return core, out

    # reference :240-267
    def get_uncertainty_explanation(self, query_text: str) -> Dict[str, Any]:
        found = self.detect_uncertainty(query_text)
        info = {
            "original_query": query_text,
            "has_uncertainty": found["has_uncertainty"],
            "processed_query": found["clean_text"],
            "uncertainty_analysis": found,
            "processing_strategy": "none",
        }
        if found["has_uncertainty"]:
            info["processing_strategy"] = "unspecified_priority"
            info["strategy_description"] = (
                f"检测到不确定性表达 {found['matched_patterns']}，"
                f"优先匹配ICD中包含'未特指'、'其他'等不确定性描述的编码"
            )
        return info
